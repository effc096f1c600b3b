// Training step of the denoiser (reference tld/train.py:160-170: forward, MSE, backward; the loss and the optimiser
// stay on the caller's side): tld_train_forward keeps the activations the backward needs, tld_train_backward turns
// d(loss)/d(pred) into the fp32 gradient of every parameter in the reference's state_dict layout.
//
// Tensor-core work: every dgrad / wgrad is the tcgen05 GEMM of gemm_tcgen05.cuh on bf16 operands
//   dgrad  dX[T,K]  = dY[T,N] * W[N,K]          -> launch_gemm_nn(A = dY, B = W as stored): MN-major B operand
//   wgrad  dW[N,K]  = dY^T[N,T] * X[T,K]        -> launch_gemm_mn(A = dY, B = X): MN-major operands, no transposes
// Everything else is the HBM-bound kernels of backward.cu / attention_bwd.cu plus a few tiny fp32 products for the
// 16-wide patch/out projections and the B-row conditioning path.  The residual-stream gradient stays fp32.
#include <algorithm>
#include <cstdlib>

#include "gemm_tcgen05.cuh"
#include "handle.h"

namespace tld {

int launch_cast_transpose_f32(const float* in, bf16* out, bf16* outT, int R, int C, cudaStream_t st);
int launch_colsum_f32(const float* in, float* out, int R, int C, int accumulate, cudaStream_t st);
int launch_colsum_bf16(const bf16* in, float* out, int R, int C, int accumulate, cudaStream_t st);
int launch_layernorm_bwd(const float* dy, const float* x, const float* gamma, float* dx, float* dgamma, float* dbeta,
                         int rows, int D, cudaStream_t st);
int launch_dwconv_gelu_bwd(const bf16* hid, const bf16* dg, const float* w9, const float* bias, bf16* du_tmp, bf16* dhid,
                           float* dw9, float* db, int B, int G, int C, cudaStream_t st);
int launch_xattn_bwd(const bf16* q, const float* go, const float* kv0, const float* kv1, long long kv_stride, bf16* dq,
                     float* dkv0, float* dkv1, long long dkv_stride, int B, int n_tok, int D, cudaStream_t st);
int launch_self_attention_bwd(const bf16* qkv, const float* d_out, const float* x_before, const float* x_after, bf16* dqkv,
                              int B, int n_tok, int D, cudaStream_t st);

// ------------------------------------------------------------------------------------------ small fp32 helpers
// C[i][j] = sum_k A(i,k) * B(k,j), generic strides, 16x16 tiles, optional split over k (partial buffers, then summed)
__global__ void __launch_bounds__(256) small_gemm_kernel(const float* __restrict__ A, long long sa_i, long long sa_k,
                                                         const float* __restrict__ B, long long sb_k, long long sb_j,
                                                         float* __restrict__ C, int M, int N, int K, int k_per_split) {
  __shared__ float sA[16][17], sB[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i = blockIdx.y * 16 + ty, j = blockIdx.x * 16 + tx;
  const int k0 = blockIdx.z * k_per_split, k1 = min(K, k0 + k_per_split);
  float acc = 0.f;
  for (int kk = k0; kk < k1; kk += 16) {
    const int ia = blockIdx.y * 16 + ty, ka = kk + tx;
    sA[ty][tx] = (ia < M && ka < k1) ? A[ia * sa_i + ka * sa_k] : 0.f;
    const int kb = kk + ty, jb = blockIdx.x * 16 + tx;
    sB[ty][tx] = (kb < k1 && jb < N) ? B[kb * sb_k + jb * sb_j] : 0.f;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; ++q) acc += sA[ty][q] * sB[q][tx];
    __syncthreads();
  }
  if (i < M && j < N) C[((size_t)blockIdx.z * M + i) * N + j] = acc;
}
__global__ void __launch_bounds__(256) sum_splits_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                         long long n, int splits) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float t = 0.f;
  for (int s = 0; s < splits; ++s) t += partial[(size_t)s * n + i];
  out[i] = t;
}
// d_pred[B,C,H,W] -> d_tok[T, pd] (transpose of the unpatchify of denoiser.py:47-52)
__global__ void __launch_bounds__(256) patch_grad_kernel(const float* __restrict__ dp, float* __restrict__ dt, int B, int C,
                                                         int img, int patch) {
  const int g = img / patch, N = g * g, pd = C * patch * patch;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)B * N * pd) return;
  const int o = int(i % pd);
  const long long tok = i / pd;
  const int n = int(tok % N), b = int(tok / N), gy = n / g, gx = n % g;
  const int c = o / (patch * patch), p1 = (o / patch) % patch, p2 = o % patch;
  dt[i] = dp[((size_t)b * C + c) * img * img + (size_t)(gy * patch + p1) * img + gx * patch + p2];
}
// out[n,d] = sum_b in[b,n,d]
__global__ void __launch_bounds__(256) batch_sum_kernel(const float* __restrict__ in, float* __restrict__ out, int B,
                                                        long long nd) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= nd) return;
  float t = 0.f;
  for (int b = 0; b < B; ++b) t += in[(size_t)b * nd + i];
  out[i] = t;
}
// LayerNorm(pd <= 64) backward per token: dx = dLN(dy; x, gamma); dgx = dy * xhat (column-summed afterwards)
__global__ void __launch_bounds__(256) ln_small_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           const float* __restrict__ gamma, float* __restrict__ dx,
                                                           float* __restrict__ dgx, long long rows, int pd) {
  const long long r = (long long)blockIdx.x * 256 + threadIdx.x;
  if (r >= rows) return;
  const float* xr = x + r * pd;
  const float* dr = dy + r * pd;
  float mu = 0.f;
  for (int i = 0; i < pd; ++i) mu += xr[i];
  mu /= pd;
  float var = 0.f;
  for (int i = 0; i < pd; ++i) var += (xr[i] - mu) * (xr[i] - mu);
  const float rstd = rsqrtf(var / pd + 1e-5f);
  float mg = 0.f, mgx = 0.f;
  for (int i = 0; i < pd; ++i) {
    const float xh = (xr[i] - mu) * rstd, gg = dr[i] * gamma[i];
    mg += gg;
    mgx += gg * xh;
  }
  mg /= pd;
  mgx /= pd;
  for (int i = 0; i < pd; ++i) {
    const float xh = (xr[i] - mu) * rstd, gg = dr[i] * gamma[i];
    dx[r * pd + i] = rstd * (gg - mg - xh * mgx);
    dgx[r * pd + i] = dr[i] * xh;
  }
}
// d_a = d_h * gelu'(a)   (exact erf GELU)
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ a,
                                                       float* __restrict__ da, long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float u = a[i];
  da[i] = dh[i] * (0.5f * (1.0f + erff(u * 0.70710678118654752f)) + u * 0.3989422804014327f * __expf(-0.5f * u * u));
}
// fp32 [R, C] -> bf16 [Rpad, C] (rows >= R zero) and bf16 transposed [C, Rpad]
__global__ void __launch_bounds__(256) cast_pad_kernel(const float* __restrict__ in, bf16* __restrict__ out,
                                                       bf16* __restrict__ outT, int R, int Rpad, int C) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)Rpad * C) return;
  const int r = int(i / C), c = int(i % C);
  const bf16 v = __float2bfloat16(r < R ? in[(size_t)r * C + c] : 0.f);
  if (out) out[i] = v;
  if (outT) outT[(size_t)c * Rpad + r] = v;
}
__global__ void __launch_bounds__(256) bf16_pad_transpose_kernel(const bf16* __restrict__ in, bf16* __restrict__ outT, int R,
                                                                 int Rpad, int C) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)Rpad * C) return;
  const int r = int(i / C), c = int(i % C);
  outT[(size_t)c * Rpad + r] = r < R ? in[(size_t)r * C + c] : __float2bfloat16(0.f);
}
__global__ void __launch_bounds__(256) transpose_f32_small_kernel(const float* __restrict__ s, float* __restrict__ d, int rows,
                                                                  int cols) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;  // d[c][r] = s[r][c]
  if (i < (long long)rows * cols) d[(size_t)(i % cols) * rows + i / cols] = s[i];
}

// C[M,N] (row-major) = sum_k A(i,k) B(k,j)
static int small_gemm(const float* A, long long sa_i, long long sa_k, const float* B, long long sb_k, long long sb_j,
                      float* C, int M, int N, int K, cudaStream_t st) {
  int splits = 1;
  if (K >= 2048) splits = (K + 511) / 512;
  if (splits > 64) splits = 64;
  int kps = (K + splits - 1) / splits;
  kps = ((kps + 15) / 16) * 16;
  splits = (K + kps - 1) / kps;
  float* dst = C;
  if (splits > 1) {
    dst = device_scratch(SCR_SPLIT_K, (size_t)splits * M * N);
    if (!dst) return 1;
  }
  small_gemm_kernel<<<dim3((N + 15) / 16, (M + 15) / 16, splits), 256, 0, st>>>(A, sa_i, sa_k, B, sb_k, sb_j, dst, M, N, K, kps);
  TLD_CUDA_OK(cudaGetLastError());
  if (splits > 1) {
    const long long n = (long long)M * N;
    sum_splits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, C, n, splits);
    TLD_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

template <typename T>
static int talloc(tld_denoiser* h, T** p, long long n) {
  void* q = nullptr;
  TLD_CUDA_OK(cudaMalloc(&q, (size_t)(n > 0 ? n : 1) * sizeof(T)));
  *p = reinterpret_cast<T*>(q);
  h->train_allocs.push_back(q);
  return 0;
}

static float* G(tld_denoiser* h, const std::string& key) { return h->grads.at(key).first; }

static int ensure_train(tld_denoiser* h, int B) {
  const int D = h->D, L = h->L, H4 = h->H4;
  if (h->grad_arena == nullptr) {
    long long total = 0;
    for (auto& kv : h->slots) total += kv.second.numel;
    TLD_CUDA_OK(cudaMalloc(&h->grad_arena, (size_t)total * sizeof(float)));
    TLD_CUDA_OK(cudaMemset(h->grad_arena, 0, (size_t)total * sizeof(float)));   // buffer slots are never written
    h->grad_elems = total;
    h->ev_grad.resize(L + 1);
    for (auto& e : h->ev_grad) TLD_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // the L kv_linear gradients must be contiguous (one wgrad GEMM over the concatenated [L*2D, D] weight)
    long long off = 0;
    for (int l = 0; l < L; ++l) {
      const std::string key = "denoiser_trans_block.decoder_blocks." + std::to_string(l) + ".cross_attention.kv_linear.weight";
      h->grads[key] = {h->grad_arena + off, 2LL * D * D};
      off += 2LL * D * D;
    }
    for (auto& kv : h->slots) {
      if (h->grads.count(kv.first)) continue;
      h->grads[kv.first] = {h->grad_arena + off, kv.second.numel};
      off += kv.second.numel;
    }
  }
  if (B <= h->train_batch) return 0;
  TLD_CUDA_OK(cudaDeviceSynchronize());
  for (void* p : h->train_allocs) cudaFree(p);
  h->train_allocs.clear();
  const long long T = (long long)B * h->N;
  const int R8 = ((2 * B + 63) / 64) * 64;
  h->tl.resize(L);
  for (int l = 0; l < L; ++l) {
    auto& t = h->tl[l];
    if (talloc(h, &t.xs0, T * D) || talloc(h, &t.xs1, T * D) || talloc(h, &t.xs2, T * D) || talloc(h, &t.qkv, T * 3 * D) ||
        talloc(h, &t.hid, T * H4) || talloc(h, &t.hid2, T * H4) || talloc(h, &t.xn0, T * D) || talloc(h, &t.xn1, T * D) ||
        talloc(h, &t.xn2, T * D))
      return 1;
  }
  const long long kvs = 2LL * L * D;
  if (talloc(h, &h->t_dx, T * D) || talloc(h, &h->t_dxn, T * D) || talloc(h, &h->t_a, T * D) ||
      talloc(h, &h->t_big, T * H4) || talloc(h, &h->t_big2, T * H4) ||
      talloc(h, &h->t_xnT, 2LL * 9 * H4) /* [9, H4] fp32 scratch of the depthwise tap gradients */ || talloc(h, &h->t_q, T * D) || talloc(h, &h->t_dkv, (long long)R8 * kvs) ||
      talloc(h, &h->t_cond_pre, 2LL * B * D) || talloc(h, &h->t_cond_h1, 1LL * B * D) || talloc(h, &h->t_cond_a1, 1LL * B * D) ||
      talloc(h, &h->t_cond_emb, 1LL * B * h->E))
    return 1;
  // t_small: patch-sized intermediates + conditioning scratch
  const long long small = 6 * T * h->pd + T * D + 8LL * R8 * D + 1LL * B * h->Te + 4LL * R8 * kvs;
  if (talloc(h, &h->t_small, small)) return 1;
  h->train_batch = B;
  return 0;
}

}  // namespace tld

extern "C" {

// pred[B,C,H,W] = Denoiser.forward(x, noise_level, label), keeping the activations for tld_train_backward.
TLD_API int tld_train_forward(tld_denoiser* h, const float* x, const float* noise_level, const float* label, float* out,
                              int batch, void* stream) {
  TLD_CHECK(h && x && noise_level && label && out && batch > 0, "tld_train_forward: bad argument");
  TLD_CHECK(tld_denoiser_missing_params(h) == 0, "tld_train_forward: parameters missing");
  TLD_CHECK(h->N <= 256 || h->N % 256 == 0, "tld_train_forward: the attention backward needs <= 256 tokens per sample or a multiple of 256");
  TLD_CHECK(h->D % 128 == 0, "tld_train_forward: the training kernels need embed_dim % 128 == 0");
  TLD_CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int D = h->D, H4 = h->H4, N = h->N, L = h->L, B = batch;
  const int T = B * N;
  if (tld_internal_ensure(h, B, 2 * B)) return 1;
  if (ensure_train(h, B)) return 1;
  h->train_serial = ++h->fwd_serial;
  const long long kvs = 2LL * L * D;
  float* s_label = h->t_small;  // [B, Te] copy of the labels for the label_proj wgrad
  TLD_CUDA_OK(cudaMemcpyAsync(s_label, label, sizeof(float) * (size_t)B * h->Te, cudaMemcpyDeviceToDevice, st));
  CondSave cs{h->t_cond_emb, h->t_cond_a1, h->t_cond_h1, h->t_cond_pre};
  if (launch_cond_noise(noise_level, B, h->E, D, h->cond, h->ycond, h->cond_scratch, st, &cs)) return 1;
  if (launch_cond_label(label, B, B, h->Te, D, h->cond, h->ycond + (size_t)B * D, h->cond_scratch, st,
                        h->t_cond_pre + (size_t)B * D))
    return 1;
  if (launch_gemm(EPI_F32, h->ycond, D, h->wkv_all, D, 2 * B, int(kvs), D, h->kv, int(kvs), nullptr, nullptr, st)) return 1;
  float* sm = h->t_small + (size_t)B * h->Te;
  EmbedSave es{sm, sm + (size_t)T * h->pd, sm + 2 * (size_t)T * h->pd, sm + 6 * (size_t)T * h->pd};
  if (launch_embed(x, B, B, h->C, h->img, h->patch, D, h->emb, h->x_res, st, &es)) return 1;
  const size_t xbytes = sizeof(float) * (size_t)T * D;
  for (int l = 0; l < L; ++l) {
    const auto& ly = h->layers[l];
    auto& t = h->tl[l];
    TLD_CUDA_OK(cudaMemcpyAsync(t.xs0, h->x_res, xbytes, cudaMemcpyDeviceToDevice, st));
    if (launch_layernorm_bf16(h->x_res, ly.ln1w, ly.ln1b, t.xn0, T, D, st)) return 1;   // kept: A operand of the qkv wgrad
    if (launch_gemm(EPI_BF16, t.xn0, D, ly.wqkv, D, T, 3 * D, D, t.qkv, 3 * D, nullptr, nullptr, st)) return 1;
    if (launch_self_attention(t.qkv, h->x_res, B, N, D, st, 0)) return 1;
    TLD_CUDA_OK(cudaMemcpyAsync(t.xs1, h->x_res, xbytes, cudaMemcpyDeviceToDevice, st));
    if (launch_layernorm_bf16(h->x_res, ly.ln2w, ly.ln2b, t.xn1, T, D, st)) return 1;
    XattnArgs xa{h->kv + (size_t)l * 2 * D, h->kv + (size_t)B * kvs + (size_t)l * 2 * D, kvs, kvs, nullptr, N, D};
    if (launch_gemm(EPI_XATTN_RESID_F32, t.xn1, D, ly.wq, D, T, D, D, h->x_res, D, nullptr, &xa, st)) return 1;
    TLD_CUDA_OK(cudaMemcpyAsync(t.xs2, h->x_res, xbytes, cudaMemcpyDeviceToDevice, st));
    if (launch_layernorm_bf16(h->x_res, ly.ln3w, ly.ln3b, t.xn2, T, D, st)) return 1;
    if (h->G == 16 && H4 % 256 == 0) {   // up-projection + depthwise conv + GELU in one kernel, hidden tensor stored as well
      if (launch_gemm_up_dwconv_gelu(t.xn2, D, ly.wup, D, T, H4, D, ly.bup, nullptr, nullptr, 0, 1e-5f, ly.dww9, ly.dwb, t.hid2, st,
                                     t.hid))
        return 1;
    } else {
      if (launch_gemm(EPI_BIAS_BF16, t.xn2, D, ly.wup, D, T, H4, D, t.hid, H4, ly.bup, nullptr, st)) return 1;
      if (launch_dwconv_gelu(t.hid, ly.dww9, ly.dwb, t.hid2, B, h->G, H4, st)) return 1;
    }
    if (launch_gemm(EPI_BIAS_RESID_F32, t.hid2, H4, ly.wdown, H4, T, D, H4, h->x_res, D, ly.bdown, nullptr, st)) return 1;
  }
  return launch_outproj(h->x_res, h->out_w, h->out_b, out, B, h->C, h->img, h->patch, D, st);
}

// d(loss)/d(pred) [B,C,H,W] -> gradients of all parameters (readable with tld_train_get_grad)
TLD_API int tld_train_backward(tld_denoiser* h, const float* d_pred, int batch, void* stream) {
  TLD_CHECK(h && d_pred && batch > 0 && batch <= h->train_batch, "tld_train_backward: call tld_train_forward with this batch first");
  TLD_CHECK(h->train_serial == h->fwd_serial,
            "tld_train_backward: the saved activations were overwritten by a later forward on this handle (one forward per "
            "backward: the activation buffers are per handle, not per autograd node)");
  TLD_CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int D = h->D, H4 = h->H4, N = h->N, L = h->L, B = batch, pd = h->pd, E = h->E, Te = h->Te;
  const int T = B * N;
  const long long kvs = 2LL * L * D;
  const int R8 = ((2 * B + 63) / 64) * 64;
  const std::string tb = "denoiser_trans_block.";
  // scratch carve-up (matches tld_train_forward)
  float* s_label = h->t_small;
  float* sm = h->t_small + (size_t)B * Te;
  float *s_u = sm, *s_c16 = sm + (size_t)T * pd, *s_t16 = sm + 2 * (size_t)T * pd, *s_dtok = sm + 3 * (size_t)T * pd;
  float *s_dt16 = sm + 4 * (size_t)T * pd, *s_dc16 = sm + 5 * (size_t)T * pd, *s_e = sm + 6 * (size_t)T * pd;
  float* s_cond = s_e + (size_t)T * D;                 // 8*R8*D floats
  float *c_dy = s_cond, *c_dpre = s_cond + (size_t)R8 * D, *c_dh1 = s_cond + 2 * (size_t)R8 * D, *c_da1 = s_cond + 3 * (size_t)R8 * D;
  bf16* c_dkvb = reinterpret_cast<bf16*>(s_cond + 8 * (size_t)R8 * D);   // [R8, kvs] bf16
  bf16* c_dkvT = c_dkvb + (size_t)R8 * kvs;                               // [kvs, R8]
  bf16* c_ycT = c_dkvT + (size_t)R8 * kvs;                                // [D, R8]
  auto blocks = [](long long n) { return (unsigned)((n + 255) / 256); };

  // ---- output projection (denoiser.py:72): pred_tok = x W_out^T + b
  patch_grad_kernel<<<blocks((long long)T * pd), 256, 0, st>>>(d_pred, s_dtok, B, h->C, h->img, h->patch);
  TLD_CUDA_OK(cudaGetLastError());
  if (launch_colsum_f32(s_dtok, G(h, tb + "out_proj.0.bias"), T, pd, 0, st)) return 1;
  if (small_gemm(s_dtok, 1, pd, h->x_res, D, 1, G(h, tb + "out_proj.0.weight"), pd, D, T, st)) return 1;   // dW = dtok^T x
  if (small_gemm(s_dtok, pd, 1, h->out_w, D, 1, h->t_dx, T, D, pd, st)) return 1;                           // dx = dtok W
  TLD_CUDA_OK(cudaMemsetAsync(h->t_dkv, 0, sizeof(float) * (size_t)R8 * kvs, st));

  for (int l = L - 1; l >= 0; --l) {
    const auto& ly = h->layers[l];
    auto& t = h->tl[l];
    const std::string b = tb + "decoder_blocks." + std::to_string(l) + ".";
    // ================= MLPSepConv: x3 = x2 + conv1x1(gelu(dwconv(conv1x1(LN3 x2)))) =================
    if (launch_cast_transpose_f32(h->t_dx, h->t_a, nullptr, T, D, st)) return 1;                     // dy -> bf16
    if (launch_colsum_f32(h->t_dx, G(h, b + "mlp.mlp.3.bias"), T, D, 0, st)) return 1;
    if (launch_gemm_nn(EPI_BF16, h->t_a, D, ly.wdown, H4, T, H4, D, h->t_big, H4, st)) return 1;   // d_hid2 = dy W_down
    if (launch_gemm_mn(EPI_F32, h->t_a, D, t.hid2, H4, D, H4, T, G(h, b + "mlp.mlp.3.weight"), H4, st)) return 1;  // dy^T hid2
    float* dw9 = reinterpret_cast<float*>(h->t_xnT);  // [9, H4] scratch
    if (launch_dwconv_gelu_bwd(t.hid, h->t_big, ly.dww9, ly.dwb, h->t_big2, h->t_big, dw9, G(h, b + "mlp.mlp.1.bias"), B, h->G, H4,
                               st))
      return 1;
    transpose_f32_small_kernel<<<blocks(9LL * H4), 256, 0, st>>>(dw9, G(h, b + "mlp.mlp.1.weight"), 9, H4);  // [9,C] -> [C,9]
    TLD_CUDA_OK(cudaGetLastError());
    if (launch_colsum_bf16(h->t_big, G(h, b + "mlp.mlp.0.bias"), T, H4, 0, st)) return 1;
    if (launch_gemm_mn(EPI_F32, h->t_big, H4, t.xn2, D, H4, D, T, G(h, b + "mlp.mlp.0.weight"), D, st)) return 1;  // d_hid^T LN3(x2)
    if (launch_gemm_nn(EPI_F32, h->t_big, H4, ly.wup, D, T, D, H4, h->t_dxn, D, st)) return 1;      // d LN3 out = d_hid W_up
    if (launch_layernorm_bwd(h->t_dxn, t.xs2, ly.ln3w, h->t_dx, G(h, b + "norm3.weight"), G(h, b + "norm3.bias"), T, D, st)) return 1;
    // ================= cross-attention: x2 = x1 + CA(LN2 x1, y) =================
    if (launch_gemm(EPI_BF16, t.xn1, D, ly.wq, D, T, D, D, h->t_q, D, nullptr, nullptr, st)) return 1;               // recompute q
    if (launch_xattn_bwd(h->t_q, h->t_dx, h->kv + (size_t)l * 2 * D, h->kv + (size_t)B * kvs + (size_t)l * 2 * D, kvs, h->t_a,
                         h->t_dkv + (size_t)l * 2 * D, h->t_dkv + (size_t)B * kvs + (size_t)l * 2 * D, kvs, B, N, D, st))
      return 1;
    if (launch_gemm_mn(EPI_F32, h->t_a, D, t.xn1, D, D, D, T, G(h, b + "cross_attention.q_linear.weight"), D, st)) return 1;  // dq^T LN2(x1)
    if (launch_gemm_nn(EPI_F32, h->t_a, D, ly.wq, D, T, D, D, h->t_dxn, D, st)) return 1;           // d LN2 out = dq W_q
    if (launch_layernorm_bwd(h->t_dxn, t.xs1, ly.ln2w, h->t_dx, G(h, b + "norm2.weight"), G(h, b + "norm2.bias"), T, D, st)) return 1;
    // ================= self-attention: x1 = x0 + Attn(qkv(LN1 x0)) =================
    if (launch_self_attention_bwd(t.qkv, h->t_dx, t.xs0, t.xs1, h->t_big, B, N, D, st)) return 1;   // dqkv [T,3D]
    if (launch_gemm_mn(EPI_F32, h->t_big, 3 * D, t.xn0, D, 3 * D, D, T, G(h, b + "self_attention.qkv_linear.weight"), D, st))
      return 1;                                                                                     // dqkv^T xn
    if (launch_gemm_nn(EPI_F32, h->t_big, 3 * D, ly.wqkv, D, T, D, 3 * D, h->t_dxn, D, st)) return 1;   // d LN1 out = dqkv W_qkv
    if (launch_layernorm_bwd(h->t_dxn, t.xs0, ly.ln1w, h->t_dx, G(h, b + "norm1.weight"), G(h, b + "norm1.bias"), T, D, st)) return 1;
    TLD_CUDA_OK(cudaEventRecord(h->ev_grad[l], st));   // this layer's gradients (all but kv_linear) are final
  }

  // ---- patch embedding (denoiser.py:34-45,75-77): tokens = LN_D(W3 LN_pd(conv(u)) + b3) + pos
  batch_sum_kernel<<<blocks((long long)N * D), 256, 0, st>>>(h->t_dx, G(h, tb + "pos_embed.weight"), B, (long long)N * D);
  TLD_CUDA_OK(cudaGetLastError());
  TLD_CUDA_OK(cudaMemsetAsync(h->t_dxn, 0, sizeof(float) * (size_t)T * D, st));
  if (launch_layernorm_bwd(h->t_dx, s_e, h->emb.ln2_w, h->t_dxn, G(h, tb + "patchify_and_embed.4.weight"),
                           G(h, tb + "patchify_and_embed.4.bias"), T, D, st))
    return 1;                                                                                          // t_dxn = d_e
  if (launch_colsum_f32(h->t_dxn, G(h, tb + "patchify_and_embed.3.bias"), T, D, 0, st)) return 1;
  if (small_gemm(h->t_dxn, 1, D, s_t16, pd, 1, G(h, tb + "patchify_and_embed.3.weight"), D, pd, T, st)) return 1;   // dW3 [D,pd]
  if (small_gemm(h->t_dxn, D, 1, h->emb.lin_wT, 1, D, s_dt16, T, pd, D, st)) return 1;   // d_t16 = d_e W3 ; W3[d][i] = lin_wT[i][d]
  ln_small_bwd_kernel<<<blocks(T), 256, 0, st>>>(s_dt16, s_c16, h->emb.ln1_w, s_dc16, s_dtok /*dgx*/, T, pd);
  TLD_CUDA_OK(cudaGetLastError());
  if (launch_colsum_f32(s_dtok, G(h, tb + "patchify_and_embed.2.weight"), T, pd, 0, st)) return 1;
  if (launch_colsum_f32(s_dt16, G(h, tb + "patchify_and_embed.2.bias"), T, pd, 0, st)) return 1;
  if (launch_colsum_f32(s_dc16, G(h, tb + "patchify_and_embed.0.bias"), T, pd, 0, st)) return 1;
  if (small_gemm(s_dc16, 1, pd, s_u, pd, 1, G(h, tb + "patchify_and_embed.0.weight"), pd, pd, T, st)) return 1;     // dWconv [pd,pd]

  // ---- conditioning path (denoiser.py:117-122): rows [0,B) noise tokens, [B,2B) label tokens
  cast_pad_kernel<<<blocks((long long)R8 * kvs), 256, 0, st>>>(h->t_dkv, c_dkvb, c_dkvT, 2 * B, R8, int(kvs));
  TLD_CUDA_OK(cudaGetLastError());
  if (launch_gemm_nn(EPI_F32, c_dkvb, int(kvs), h->wkv_all, D, R8, D, int(kvs), c_dy, D, st)) return 1;  // d y = dkv W_kv
  bf16_pad_transpose_kernel<<<blocks((long long)R8 * D), 256, 0, st>>>(h->ycond, c_ycT, 2 * B, R8, D);
  TLD_CUDA_OK(cudaGetLastError());
  if (launch_gemm(EPI_F32, c_dkvT, R8, c_ycT, R8, int(kvs), D, R8, G(h, tb + "decoder_blocks.0.cross_attention.kv_linear.weight"), D,
                  nullptr, nullptr, st))
    return 1;                                                                                          // all layers' dWkv
  TLD_CUDA_OK(cudaMemsetAsync(c_dpre, 0, sizeof(float) * (size_t)R8 * D, st));
  if (launch_layernorm_bwd(c_dy, h->t_cond_pre, h->cond.ln_w, c_dpre, G(h, "norm.weight"), G(h, "norm.bias"), 2 * B, D, st)) return 1;
  const float* dpre_n = c_dpre;
  const float* dpre_l = c_dpre + (size_t)B * D;
  if (launch_colsum_f32(dpre_l, G(h, "label_proj.bias"), B, D, 0, st)) return 1;
  if (small_gemm(dpre_l, 1, D, s_label, Te, 1, G(h, "label_proj.weight"), D, Te, B, st)) return 1;
  if (launch_colsum_f32(dpre_n, G(h, "fourier_feats.3.bias"), B, D, 0, st)) return 1;
  if (small_gemm(dpre_n, 1, D, h->t_cond_h1, D, 1, G(h, "fourier_feats.3.weight"), D, D, B, st)) return 1;
  if (small_gemm(dpre_n, D, 1, h->cond.w2, D, 1, c_dh1, B, D, D, st)) return 1;                      // d_h1 = d_pre W2
  gelu_bwd_kernel<<<blocks((long long)B * D), 256, 0, st>>>(c_dh1, h->t_cond_a1, c_da1, (long long)B * D);
  TLD_CUDA_OK(cudaGetLastError());
  if (launch_colsum_f32(c_da1, G(h, "fourier_feats.1.bias"), B, D, 0, st)) return 1;
  if (small_gemm(c_da1, 1, D, h->t_cond_emb, E, 1, G(h, "fourier_feats.1.weight"), D, E, B, st)) return 1;
  // the frequency buffer is not a parameter
  TLD_CUDA_OK(cudaMemsetAsync(G(h, "fourier_feats.0.angular_speeds"), 0, sizeof(float) * (size_t)(E / 2), st));
  TLD_CUDA_OK(cudaEventRecord(h->ev_grad[L], st));
  return 0;
}

// create the gradient arena (and its layout) without running a forward: lets an optimiser lay its flat parameter / moment
// arenas out like the gradients before the first step
TLD_API int tld_train_prepare(tld_denoiser* h) {
  TLD_CHECK(h, "tld_train_prepare: null handle");
  TLD_CUDA_OK(cudaSetDevice(h->device));
  return ensure_train(h, 0);
}

// copy the gradient of `key` (reference state_dict key / layout) to dst (device, fp32)
TLD_API int tld_train_get_grad(tld_denoiser* h, const char* key, float* dst, int64_t numel, void* stream) {
  TLD_CHECK(h && key && dst, "tld_train_get_grad: null argument");
  auto it = h->grads.find(key);
  if (it == h->grads.end()) return fail(std::string("tld_train_get_grad: unknown key ") + key);
  TLD_CHECK(numel == it->second.second, std::string("tld_train_get_grad: size mismatch for ") + key);
  TLD_CUDA_OK(cudaMemcpyAsync(dst, it->second.first, sizeof(float) * (size_t)numel, cudaMemcpyDeviceToDevice,
                              reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

// ---- data-parallel support: the gradient arena as ranges that become final at known points of the backward, so that the
// caller's all-reduce of layer l can run while layers l-1 .. 0 are still being differentiated (tld/train.py:169:
// accelerator.backward = DDP's bucketed, overlapped all-reduce).
//   segment l in [0, L): every gradient of decoder block l except its kv_linear weight;  segment L: the concatenated kv_linear
//   gradients of all blocks;  segment L+1: everything else (patch embedding, positions, out-projection, conditioning MLP).
// out[2*i] = element offset into the arena, out[2*i+1] = elements.  *arena receives the device pointer (fp32).
TLD_API int tld_train_grad_layout(tld_denoiser* h, float** arena, int64_t* total, int64_t* out, int n_segments) {
  TLD_CHECK(h && arena && total && out, "tld_train_grad_layout: null argument");
  TLD_CHECK(h->grad_arena != nullptr, "tld_train_grad_layout: call tld_train_forward first");
  const int L = h->L;
  TLD_CHECK(n_segments == L + 2, "tld_train_grad_layout: n_segments must be n_layers + 2");
  *arena = h->grad_arena;
  *total = h->grad_elems;
  const long long kv_elems = 2LL * h->D * h->D * L;
  long long tail_lo = h->grad_elems, tail_sum = 0;
  std::vector<long long> lo(L, h->grad_elems), hi(L, 0), sum(L, 0);
  const std::string pre = "denoiser_trans_block.decoder_blocks.";
  for (auto& kv : h->grads) {
    const long long off = kv.second.first - h->grad_arena, n = kv.second.second;
    if (off < kv_elems) continue;  // the kv block
    if (kv.first.compare(0, pre.size(), pre) == 0) {
      const int l = std::atoi(kv.first.c_str() + pre.size());
      TLD_CHECK(l >= 0 && l < L, "tld_train_grad_layout: bad layer index in key");
      lo[l] = std::min(lo[l], off);
      hi[l] = std::max(hi[l], off + n);
      sum[l] += n;
    } else {
      tail_lo = std::min(tail_lo, off);
      tail_sum += n;
    }
  }
  for (int l = 0; l < L; ++l) {
    TLD_CHECK(hi[l] - lo[l] == sum[l], "tld_train_grad_layout: a layer's gradients are not contiguous");
    out[2 * l] = lo[l];
    out[2 * l + 1] = sum[l];
  }
  out[2 * L] = 0;
  out[2 * L + 1] = kv_elems;
  TLD_CHECK(tail_lo + tail_sum == h->grad_elems, "tld_train_grad_layout: tail gradients are not contiguous");
  out[2 * L + 2] = tail_lo;
  out[2 * L + 3] = tail_sum;
  return 0;
}

// element offset / size of one parameter's gradient inside the arena (reference state_dict key and layout)
TLD_API int tld_train_grad_offset(tld_denoiser* h, const char* key, int64_t* offset, int64_t* numel) {
  TLD_CHECK(h && key && offset && numel, "tld_train_grad_offset: null argument");
  auto it = h->grads.find(key);
  if (it == h->grads.end()) return fail(std::string("tld_train_grad_offset: unknown key ") + key);
  *offset = it->second.first - h->grad_arena;
  *numel = it->second.second;
  return 0;
}

// make `stream` wait until segment `segment` of the running backward is final (segment < n_layers: that layer; otherwise the
// whole backward).  Pure stream ordering, the host does not block.
TLD_API int tld_train_wait_grad(tld_denoiser* h, int segment, void* stream) {
  TLD_CHECK(h && !h->ev_grad.empty(), "tld_train_wait_grad: no backward has been recorded");
  const int L = h->L;
  const int idx = (segment >= 0 && segment < L) ? segment : L;
  TLD_CUDA_OK(cudaStreamWaitEvent(reinterpret_cast<cudaStream_t>(stream), h->ev_grad[idx], 0));
  return 0;
}

}  // extern "C"
