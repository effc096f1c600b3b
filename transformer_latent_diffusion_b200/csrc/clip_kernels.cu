// Row-wise kernels of the CLIP text tower: the reference encodes the prompt with
// `clip_model.encode_text(clip.tokenize(prompt))` (tld/diffusion.py:136-140,160,177; openai/CLIP model.py: token + positional
// embedding -> 12 pre-LN residual attention blocks with a causal mask -> ln_final -> features at the EOT token -> text_projection).
// The GEMMs (in_proj, out_proj, c_fc, c_proj) are the tcgen05 GEMM of gemm_tcgen05.cuh with its bias / bias+residual epilogues
// and the LayerNorms are layernorm_bf16_kernel; this file holds what is left - all of it tiny (77 tokens per prompt):
//   clip_embed_kernel            x[b,t,:] = token_embedding[ids[b,t]] + positional_embedding[t]          (fp32 residual stream)
//   clip_causal_attention_kernel softmax(q k^T / 8 + causal mask) v per (prompt, head), head_dim 64, <= 128 tokens
//   quick_gelu_kernel            x * sigmoid(1.702 x)  (openai/CLIP QuickGELU)
//   clip_final_kernel            ln_final(x[b, eot_b]) @ text_projection                                   (fp32)
#include <math.h>

#include "../../include/tld_b200.h"
#include "common.h"

namespace tld {

__global__ void __launch_bounds__(256) clip_embed_kernel(const long long* __restrict__ ids, const float* __restrict__ tok,
                                                         const float* __restrict__ pos, float* __restrict__ x, int n_ctx,
                                                         int D, long long rows, int vocab) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;   // one thread = 4 columns of one row
  const int d4 = D / 4;
  if (i >= rows * d4) return;
  const long long r = i / d4;
  const int c = int(i % d4) * 4;
  long long id = ids[r];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float4 a = *reinterpret_cast<const float4*>(tok + id * D + c);
  const float4 p = *reinterpret_cast<const float4*>(pos + (size_t)(r % n_ctx) * D + c);
  *reinterpret_cast<float4*>(x + r * D + c) = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
}

// one CTA = one (prompt, head); K and V of the head as fp32 in shared memory (row pitch 65: conflict-free column walks);
// warp w takes queries w, w + 8, ...: lane = key for the scores (keys j <= t only: causal), lane = 2 output dims for P V.
constexpr int CLIP_MAXT = 128;
__global__ void __launch_bounds__(256) clip_causal_attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                                    int n_ctx, int D) {
  extern __shared__ float cs[];
  float* sK = cs;                       // [n_ctx][65]
  float* sV = cs + (size_t)n_ctx * 65;  // [n_ctx][65]
  float* sP = sV + (size_t)n_ctx * 65;  // [8 warps][CLIP_MAXT]
  const int head = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ld = 3LL * D;
  const bf16* base = qkv + (size_t)b * n_ctx * ld + head * 64;
  for (int i = threadIdx.x; i < n_ctx * 64; i += 256) {
    const int t = i >> 6, d = i & 63;
    sK[t * 65 + d] = __bfloat162float(base[(size_t)t * ld + D + d]);
    sV[t * 65 + d] = __bfloat162float(base[(size_t)t * ld + 2 * D + d]);
  }
  __syncthreads();
  float* p = sP + warp * CLIP_MAXT;
  for (int t = warp; t < n_ctx; t += 8) {
    // q row in registers: lane holds dims 2*lane, 2*lane+1
    const __nv_bfloat162 q2 = *reinterpret_cast<const __nv_bfloat162*>(base + (size_t)t * ld + 2 * lane);
    const float q0 = __bfloat162float(q2.x), q1 = __bfloat162float(q2.y);
    float mx = -INFINITY;
    for (int j0 = 0; j0 <= t; j0 += 32) {
      const int j = j0 + lane;
      float s = 0.f;
#pragma unroll 8
      for (int d = 0; d < 64; d += 2) {   // q broadcast from the lane that owns the dims
        const float a0 = __shfl_sync(0xffffffffu, q0, d >> 1), a1 = __shfl_sync(0xffffffffu, q1, d >> 1);
        if (j <= t) s += a0 * sK[j * 65 + d] + a1 * sK[j * 65 + d + 1];
      }
      s *= 0.125f;
      if (j <= t) {
        p[j] = s;
        mx = fmaxf(mx, s);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncwarp();
    float sum = 0.f;
    for (int j = lane; j <= t; j += 32) {
      const float e = __expf(p[j] - mx);
      p[j] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    const float inv = 1.f / sum;
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j <= t; ++j) {
      const float pj = p[j];
      o0 += pj * sV[j * 65 + 2 * lane];
      o1 += pj * sV[j * 65 + 2 * lane + 1];
    }
    *reinterpret_cast<__nv_bfloat162*>(out + ((size_t)b * n_ctx + t) * D + head * 64 + 2 * lane) =
        __floats2bfloat162_rn(o0 * inv, o1 * inv);
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) quick_gelu_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, long long n2) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;   // bf16 pairs
  if (i >= n2) return;
  const __nv_bfloat162 v = reinterpret_cast<const __nv_bfloat162*>(in)[i];
  const float a = __bfloat162float(v.x), b = __bfloat162float(v.y);
  reinterpret_cast<__nv_bfloat162*>(out)[i] = __floats2bfloat162_rn(a / (1.f + __expf(-1.702f * a)), b / (1.f + __expf(-1.702f * b)));
}

// one CTA per prompt: LayerNorm of the EOT row (fp32), then y = xn @ proj with proj [D, P] row-major (x @ text_projection)
__global__ void __launch_bounds__(256) clip_final_kernel(const float* __restrict__ x, const long long* __restrict__ eot,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const float* __restrict__ proj, float* __restrict__ out, int n_ctx,
                                                         int D, int P) {
  extern __shared__ float fs[];   // [D] normalised row + [8] reduction scratch
  __shared__ float red[2][8];
  const int b = blockIdx.x;
  const float* xr = x + ((size_t)b * n_ctx + eot[b]) * D;
  float s = 0.f;
  for (int i = threadIdx.x; i < D; i += 256) s += xr[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s;
  __syncthreads();
  float mu = 0.f;
  for (int w = 0; w < 8; ++w) mu += red[0][w];
  mu /= D;
  float q = 0.f;
  for (int i = threadIdx.x; i < D; i += 256) q += (xr[i] - mu) * (xr[i] - mu);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = q;
  __syncthreads();
  float var = 0.f;
  for (int w = 0; w < 8; ++w) var += red[1][w];
  const float rstd = rsqrtf(var / D + 1e-5f);
  for (int i = threadIdx.x; i < D; i += 256) fs[i] = (xr[i] - mu) * rstd * gamma[i] + beta[i];
  __syncthreads();
  for (int c = threadIdx.x; c < P; c += 256) {
    float acc = 0.f;
    for (int k = 0; k < D; ++k) acc += fs[k] * proj[(size_t)k * P + c];
    out[(size_t)b * P + c] = acc;
  }
}

}  // namespace tld

using namespace tld;

extern "C" {

TLD_API int tld_clip_embed(const int64_t* ids, const float* token_embedding, const float* positional_embedding, float* x,
                           int batch, int n_ctx, int D, int vocab, void* stream) {
  TLD_CHECK(ids && token_embedding && positional_embedding && x && batch > 0 && n_ctx > 0 && D % 4 == 0, "tld_clip_embed: bad argument");
  const long long rows = (long long)batch * n_ctx;
  clip_embed_kernel<<<(unsigned)((rows * (D / 4) + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ids), token_embedding, positional_embedding, x, n_ctx, D, rows, vocab);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

TLD_API int tld_clip_causal_attention(const uint16_t* qkv, uint16_t* out, int batch, int n_ctx, int D, void* stream) {
  TLD_CHECK(qkv && out && batch > 0 && batch <= 65535 && D % 64 == 0 && n_ctx > 0 && n_ctx <= CLIP_MAXT,
            "tld_clip_causal_attention: needs embed_dim % 64 == 0 and at most 128 tokens");
  const int smem = (2 * n_ctx * 65 + 8 * CLIP_MAXT) * 4;
  static bool set = false;
  if (!set) {
    TLD_CUDA_OK(cudaFuncSetAttribute(clip_causal_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (2 * CLIP_MAXT * 65 + 8 * CLIP_MAXT) * 4));
    set = true;
  }
  clip_causal_attention_kernel<<<dim3(D / 64, batch), 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), n_ctx, D);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

TLD_API int tld_clip_quick_gelu(const uint16_t* in, uint16_t* out, long long n, void* stream) {
  TLD_CHECK(in && out && n > 0 && n % 2 == 0, "tld_clip_quick_gelu: bad argument");
  quick_gelu_kernel<<<(unsigned)((n / 2 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(in), reinterpret_cast<bf16*>(out), n / 2);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

TLD_API int tld_clip_final(const float* x, const int64_t* eot, const float* gamma, const float* beta, const float* proj, float* out,
                           int batch, int n_ctx, int D, int P, void* stream) {
  TLD_CHECK(x && eot && gamma && beta && proj && out && batch > 0 && D > 0 && P > 0, "tld_clip_final: bad argument");
  clip_final_kernel<<<batch, 256, D * 4, reinterpret_cast<cudaStream_t>(stream)>>>(x, reinterpret_cast<const long long*>(eot), gamma, beta,
                                                                                   proj, out, n_ctx, D, P);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
