// Fused Adam + EMA step (SURVEY.md §2.2 O1, §8b tld_adam_ema_step): the reference's
//   optimizer.step()                         tld/train.py:170  (torch.optim.Adam, lr 3e-4, betas (0.9, 0.999), eps 1e-8)
//   update_ema(ema_model, model, alpha)      tld/train.py:55-58,172-173
// as ONE pass over flat fp32 arenas.  HBM-bound: 4 reads + 3 writes per element (+ 1 read + 1 write with the EMA), 28-36 B
// per parameter; 101.2 M parameters -> 2.8-3.6 GB -> ~0.5 ms at the measured 6.5 TB/s (torch's multi_tensor_apply Adam
// + two foreach EMA kernels took 1.1 ms of the batch-32 step).  Grid = 148 SMs x 8 resident CTAs, grid-stride, 16-byte
// accesses.  Arithmetic follows torch.optim.Adam's single-tensor statement op for op:
//   m = lerp(m, g, 1-b1);  v = b2 v + (1-b2) g g;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps);  ema = a ema + (1-a) p
#include <math.h>

#include "common.h"
#include "../../include/tld_b200.h"

namespace tld {

struct AdamArgs {
  float beta1, beta2, one_minus_b1, one_minus_b2;
  float step_size;        // lr / (1 - beta1^t)
  float inv_bc2_sqrt;     // 1 / sqrt(1 - beta2^t)
  float eps, weight_decay;
  float ema_alpha, one_minus_alpha;
  float grad_scale;       // g *= grad_scale first (1 = none): folds a deferred gradient average into the step
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamArgs& a) {
  g *= a.grad_scale;
  if (a.weight_decay != 0.f) g = fmaf(a.weight_decay, p, g);
  m = fmaf(a.one_minus_b1, g - m, m);                 // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(a.one_minus_b2 * g, g, a.beta2 * v);       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) * a.inv_bc2_sqrt + a.eps;
  p -= a.step_size * (m / denom);                     // param.addcdiv_(exp_avg, denom, value=-step_size)
}

template <bool EMA>
__global__ void __launch_bounds__(256) adam_ema_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                       float4* __restrict__ m, float4* __restrict__ v,
                                                       float4* __restrict__ ema, long long n4, float* p_tail,
                                                       const float* g_tail, float* m_tail, float* v_tail, float* ema_tail,
                                                       int tail, AdamArgs a) {
  const long long stride = (long long)gridDim.x * 256;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
    float4 pp = p[i], mm = m[i], vv = v[i];
    const float4 gg = __ldcs(g + i);   // gradients are read exactly once
    adam_elem(pp.x, gg.x, mm.x, vv.x, a);
    adam_elem(pp.y, gg.y, mm.y, vv.y, a);
    adam_elem(pp.z, gg.z, mm.z, vv.z, a);
    adam_elem(pp.w, gg.w, mm.w, vv.w, a);
    p[i] = pp;
    m[i] = mm;
    v[i] = vv;
    if constexpr (EMA) {
      float4 e = ema[i];
      e.x = fmaf(a.ema_alpha, e.x, a.one_minus_alpha * pp.x);
      e.y = fmaf(a.ema_alpha, e.y, a.one_minus_alpha * pp.y);
      e.z = fmaf(a.ema_alpha, e.z, a.one_minus_alpha * pp.z);
      e.w = fmaf(a.ema_alpha, e.w, a.one_minus_alpha * pp.w);
      ema[i] = e;
    }
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail) {   // n % 4 trailing elements
    const int i = threadIdx.x;
    float pp = p_tail[i], mm = m_tail[i], vv = v_tail[i];
    adam_elem(pp, g_tail[i], mm, vv, a);
    p_tail[i] = pp;
    m_tail[i] = mm;
    v_tail[i] = vv;
    if constexpr (EMA) ema_tail[i] = fmaf(a.ema_alpha, ema_tail[i], a.one_minus_alpha * pp);
  }
}

}  // namespace tld

using namespace tld;

extern "C" {

TLD_API int tld_adam_ema_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, int64_t n,
                              double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                              double ema_alpha, double grad_scale, void* stream) {
  TLD_CHECK(param && grad && exp_avg && exp_avg_sq && n > 0, "tld_adam_ema_step: null argument");
  TLD_CHECK(step >= 1, "tld_adam_ema_step: step counts from 1 (the value AFTER torch's state['step'] += 1)");
  TLD_CHECK(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
              reinterpret_cast<uintptr_t>(exp_avg_sq) | reinterpret_cast<uintptr_t>(ema)) & 15) == 0,
            "tld_adam_ema_step: arenas must be 16-byte aligned");
  AdamArgs a;
  // hyper-parameters arrive as doubles (python floats) and every derived constant is formed in double before it is rounded
  // to fp32, exactly as torch does with its scalar arguments: (float)(1 - 0.999) != 1.f - (float)0.999 by 1.3e-5 relative
  a.beta1 = (float)beta1;
  a.beta2 = (float)beta2;
  a.one_minus_b1 = (float)(1.0 - beta1);
  a.one_minus_b2 = (float)(1.0 - beta2);
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);   // 1 - beta ** step
  a.step_size = (float)(lr / bc1);
  a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  a.eps = (float)eps;
  a.weight_decay = (float)weight_decay;
  a.ema_alpha = (float)ema_alpha;
  a.one_minus_alpha = (float)(1.0 - ema_alpha);
  a.grad_scale = (float)grad_scale;
  const long long n4 = n / 4;
  const int tail = int(n - 4 * n4);
  long long want = (n4 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float4 *p4 = reinterpret_cast<float4*>(param), *m4 = reinterpret_cast<float4*>(exp_avg),
         *v4 = reinterpret_cast<float4*>(exp_avg_sq), *e4 = reinterpret_cast<float4*>(ema);
  const float4* g4 = reinterpret_cast<const float4*>(grad);
  if (ema)
    adam_ema_kernel<true><<<grid, 256, 0, st>>>(p4, g4, m4, v4, e4, n4, param + 4 * n4, grad + 4 * n4, exp_avg + 4 * n4,
                                                exp_avg_sq + 4 * n4, ema + 4 * n4, tail, a);
  else
    adam_ema_kernel<false><<<grid, 256, 0, st>>>(p4, g4, m4, v4, nullptr, n4, param + 4 * n4, grad + 4 * n4, exp_avg + 4 * n4,
                                                 exp_avg_sq + 4 * n4, nullptr, tail, a);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
