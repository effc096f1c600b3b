// Backward-pass kernels of the denoiser (training step, reference tld/train.py:160-170 -> autograd through
// tld/transformer_blocks.py:135-139).  The tensor-core work of the backward pass (dgrad / wgrad) reuses the tcgen05
// GEMM of gemm_tcgen05.cuh on explicitly transposed bf16 copies (wgrad: dW[N,K] = dY^T[N,T] * X^T[K,T]^T); this file
// holds the HBM-bound pieces around it: cast+transpose, column sums (bias grads), LayerNorm backward,
// depthwise-conv/GELU backward, the 2-key cross-attention backward and the self-attention backward.
#include <math.h>

#include "common.h"

namespace tld {

__device__ __forceinline__ float bw_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// cast + transpose: in [R, C] (fp32 or bf16) -> out bf16 [R, C] (optional) and outT bf16 [C, R] (optional)
// ------------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void __launch_bounds__(256) cast_transpose_kernel(const TIn* __restrict__ in, bf16* __restrict__ out,
                                                             bf16* __restrict__ outT, int R, int C) {
  __shared__ bf16 tile[64][66];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r = r0 + ty + 4 * i, c = c0 + tx;
    bf16 v = __float2bfloat16(0.f);
    if (r < R && c < C) {
      if constexpr (sizeof(TIn) == 4) v = __float2bfloat16(in[(size_t)r * C + c]);
      else v = in[(size_t)r * C + c];
      if (out) out[(size_t)r * C + c] = v;
    }
    tile[ty + 4 * i][tx] = v;
  }
  __syncthreads();
  if (outT) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = c0 + ty + 4 * i, r = r0 + tx;
      if (r < R && c < C) outT[(size_t)c * R + r] = tile[tx][ty + 4 * i];
    }
  }
}

int launch_cast_transpose_f32(const float* in, bf16* out, bf16* outT, int R, int C, cudaStream_t st) {
  dim3 grid((C + 63) / 64, (R + 63) / 64);
  cast_transpose_kernel<float><<<grid, 256, 0, st>>>(in, out, outT, R, C);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// column sums: out[c] (+)= sum_r in[r, c]; deterministic (fixed row partition, fixed reduction order)
// ------------------------------------------------------------------------------------------------
// Two passes: partial[chunk][c] over a fixed row partition (grid = column groups x row chunks, so a [8192 x 768] input
// runs on 24 x 64 CTAs instead of 24), then a fixed-order sum over the chunks.
template <typename TIn>
__global__ void __launch_bounds__(256) colsum_kernel(const TIn* __restrict__ in, float* __restrict__ partial, int R, int C,
                                                     int rows_per_chunk) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(R, r0 + rows_per_chunk);
  float s = 0.f;
  if (c < C)
    for (int r = r0 + rl; r < r1; r += 8) {
      if constexpr (sizeof(TIn) == 4) s += in[(size_t)r * C + c];
      else s += __bfloat162float(in[(size_t)r * C + c]);
    }
  red[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    partial[(size_t)blockIdx.y * C + c] = t;
  }
}
__global__ void __launch_bounds__(256) colsum_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                            int nchunk, int C, int accumulate) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  float t = 0.f;
  for (int k = 0; k < nchunk; ++k) t += partial[(size_t)k * C + c];
  out[c] = accumulate ? out[c] + t : t;
}
template <typename TIn>
static int colsum_launch(const TIn* in, float* out, int R, int C, int accumulate, cudaStream_t st) {
  int nchunk = (R + 127) / 128;
  if (nchunk > 64) nchunk = 64;
  if (nchunk < 1) nchunk = 1;
  const int rpc = (((R + nchunk - 1) / nchunk) + 7) / 8 * 8;
  nchunk = (R + rpc - 1) / rpc;
  if (nchunk < 1) nchunk = 1;
  float* g_colsum_partial = device_scratch(SCR_COLSUM, (size_t)nchunk * C);
  if (!g_colsum_partial) return 1;
  colsum_kernel<TIn><<<dim3((C + 31) / 32, nchunk), 256, 0, st>>>(in, g_colsum_partial, R, C, rpc);
  TLD_CUDA_OK(cudaGetLastError());
  colsum_reduce_kernel<<<(C + 255) / 256, 256, 0, st>>>(g_colsum_partial, out, nchunk, C, accumulate);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}
int launch_colsum_f32(const float* in, float* out, int R, int C, int accumulate, cudaStream_t st) {
  return colsum_launch<float>(in, out, R, C, accumulate, st);
}
int launch_colsum_bf16(const bf16* in, float* out, int R, int C, int accumulate, cudaStream_t st) {
  return colsum_launch<bf16>(in, out, R, C, accumulate, st);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward (eps 1e-5, affine).  y = xhat*gamma + beta, xhat = (x - mu) * rstd.
//   dx[r,:] += rstd * (g - mean(g) - xhat * mean(g * xhat)),   g = dy * gamma
//   partial[blk][0][c] = sum_rows dy*xhat (dgamma),  partial[blk][1][c] = sum_rows dy (dbeta)
// One warp per row, 64 rows per block; a lane owns the same columns in every row, so the dgamma/dbeta partial sums
// stay in registers and are reduced across the block's 8 warps at the end (deterministic).
// ------------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ gamma, float* __restrict__ dx,
                                                            float* __restrict__ partial, int rows) {
  constexpr int D = V * 128;
  __shared__ __align__(16) float acc[2][D];  // [dgamma|dbeta][D], warps add in a fixed order (deterministic)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 dg[V], db[V], gm[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    dg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    gm[j] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * j);
  }
  const int row_base = blockIdx.x * 64;
  for (int i = 0; i < 8; ++i) {
    const int row = row_base + i * 8 + warp;
    if (row >= rows) break;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
    const float4* dyr = reinterpret_cast<const float4*>(dy + (size_t)row * D);
    float4 xv[V], dv[V];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      xv[j] = xr[lane + 32 * j];
      dv[j] = dyr[lane + 32 * j];
      s += (xv[j].x + xv[j].y) + (xv[j].z + xv[j].w);
    }
    const float mu = bw_warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float a = xv[j].x - mu, b = xv[j].y - mu, c = xv[j].z - mu, d = xv[j].w - mu;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(bw_warp_sum(q) * (1.0f / D) + 1e-5f);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      // xhat in xv, g = dy*gamma in place
      xv[j].x = (xv[j].x - mu) * rstd; xv[j].y = (xv[j].y - mu) * rstd;
      xv[j].z = (xv[j].z - mu) * rstd; xv[j].w = (xv[j].w - mu) * rstd;
      dg[j].x += dv[j].x * xv[j].x; dg[j].y += dv[j].y * xv[j].y; dg[j].z += dv[j].z * xv[j].z; dg[j].w += dv[j].w * xv[j].w;
      db[j].x += dv[j].x; db[j].y += dv[j].y; db[j].z += dv[j].z; db[j].w += dv[j].w;
      dv[j].x *= gm[j].x; dv[j].y *= gm[j].y; dv[j].z *= gm[j].z; dv[j].w *= gm[j].w;
      sg += (dv[j].x + dv[j].y) + (dv[j].z + dv[j].w);
      sgx += (dv[j].x * xv[j].x + dv[j].y * xv[j].y) + (dv[j].z * xv[j].z + dv[j].w * xv[j].w);
    }
    const float mg = bw_warp_sum(sg) * (1.0f / D), mgx = bw_warp_sum(sgx) * (1.0f / D);
    float4* dxr = reinterpret_cast<float4*>(dx + (size_t)row * D);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float4 o = dxr[lane + 32 * j];
      o.x += rstd * (dv[j].x - mg - xv[j].x * mgx);
      o.y += rstd * (dv[j].y - mg - xv[j].y * mgx);
      o.z += rstd * (dv[j].z - mg - xv[j].z * mgx);
      o.w += rstd * (dv[j].w - mg - xv[j].w * mgx);
      dxr[lane + 32 * j] = o;
    }
  }
  // block reduction of the per-warp dgamma/dbeta partials, one warp at a time
  for (int w = 0; w < 8; ++w) {
    if (warp == w) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float4* a0 = reinterpret_cast<float4*>(acc[0]) + lane + 32 * j;
        float4* a1 = reinterpret_cast<float4*>(acc[1]) + lane + 32 * j;
        if (w == 0) {
          *a0 = dg[j];
          *a1 = db[j];
        } else {
          float4 u = *a0, v = *a1;
          u.x += dg[j].x; u.y += dg[j].y; u.z += dg[j].z; u.w += dg[j].w;
          v.x += db[j].x; v.y += db[j].y; v.z += db[j].z; v.w += db[j].w;
          *a0 = u;
          *a1 = v;
        }
      }
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < 2 * D; c += 256) partial[(size_t)blockIdx.x * 2 * D + c] = acc[c / D][c % D];
}

// out[c] (+)= sum_blk partial[blk][c]   (second stage of the dgamma/dbeta reduction; C = 2*D layout [2][D])
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, int nblk, int D) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= 2 * D) return;
  float t = 0.f;
  for (int b = 0; b < nblk; ++b) t += partial[(size_t)b * 2 * D + c];
  if (c < D) dgamma[c] = t;
  else dbeta[c - D] = t;
}


int launch_layernorm_bwd(const float* dy, const float* x, const float* gamma, float* dx, float* dgamma, float* dbeta,
                         int rows, int D, cudaStream_t st) {
  TLD_CHECK(D % 128 == 0 && D >= 128 && D <= 1024, "layernorm_bwd: embed_dim must be a multiple of 128 in [128,1024]");
  const int nblk = (rows + 63) / 64;
  float* g_ln_partial = device_scratch(SCR_LN_BWD, (size_t)nblk * 2 * D);
  if (!g_ln_partial) return 1;
  switch (D / 128) {
#define LNB_CASE(V)                                                                                         \
  case V: {                                                                                                 \
    auto kern = layernorm_bwd_kernel<V>;                                                                    \
    kern<<<nblk, 256, 0, st>>>(dy, x, gamma, dx, g_ln_partial, rows);                                       \
  } break;
    LNB_CASE(1) LNB_CASE(2) LNB_CASE(3) LNB_CASE(4) LNB_CASE(5) LNB_CASE(6)
#undef LNB_CASE
    default: return fail("layernorm_bwd: embed_dim > 768 not instantiated");
  }
  TLD_CUDA_OK(cudaGetLastError());
  reduce_partials_kernel<<<(2 * D + 255) / 256, 256, 0, st>>>(g_ln_partial, dgamma, dbeta, nblk, D);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}


// ------------------------------------------------------------------------------------------------
// MLPSepConv middle backward (transformer_blocks.py:96-103):  g = gelu(u), u = dwconv3x3(h) + b  (token-major NHWC)
//   (A) du = dg * gelu'(u)                      (u recomputed from h; gelu'(u) = Phi(u) + u*phi(u))
//   (B) dh = dwconv3x3^T(du)                    (correlation with the flipped kernel, zero padding)
//   (C) dw[tap][c] = sum_{b,y,x} du[y,x,c] * h[y+dy-1, x+dx-1, c],  db[c] = sum du
// Thread = 8 channels of one position for (A)/(B) (9 x 16-byte loads from L1/L2), 2 channels x a position range for (C).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bw_unpack8(const uint4& v, float (&f)[8]) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t bw_pk2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// mode 0: out = dg * gelu'(conv(h) + bias)   (in = h, second = dg)
// mode 1: out = conv^T(in)                   (in = du, flipped taps, no bias)
__global__ void __launch_bounds__(256) dwconv_bwd_kernel(const bf16* __restrict__ in, const bf16* __restrict__ second,
                                                         const float* __restrict__ w9, const float* __restrict__ bias,
                                                         bf16* __restrict__ out, int B, int G, int C, int mode) {
  const int c8n = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * G * G * c8n) return;
  const int c0 = int(idx % c8n) * 8;
  const long long pos = idx / c8n;
  const int xq = int(pos % G), yq = int((pos / G) % G);
  const long long b = pos / ((long long)G * G);
  const size_t img = (size_t)b * G * G * C;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = (mode == 0) ? bias[c0 + j] : 0.f;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int yy = yq + dy - 1, xx = xq + dx - 1;
      if (yy < 0 || yy >= G || xx < 0 || xx >= G) continue;
      const uint4 v = *reinterpret_cast<const uint4*>(in + img + ((size_t)yy * G + xx) * C + c0);
      float f[8];
      bw_unpack8(v, f);
      // forward: tap (dy,dx) reads h[y+dy-1, x+dx-1]; transpose: du[y+dy-1, x+dx-1] contributed through tap (2-dy,2-dx)
      const int tap = (mode == 0) ? dy * 3 + dx : (2 - dy) * 3 + (2 - dx);
      const float4 wa = __ldg(reinterpret_cast<const float4*>(w9 + (size_t)tap * C + c0));
      const float4 wb = __ldg(reinterpret_cast<const float4*>(w9 + (size_t)tap * C + c0) + 1);
      acc[0] += wa.x * f[0]; acc[1] += wa.y * f[1]; acc[2] += wa.z * f[2]; acc[3] += wa.w * f[3];
      acc[4] += wb.x * f[4]; acc[5] += wb.y * f[5]; acc[6] += wb.z * f[6]; acc[7] += wb.w * f[7];
    }
  if (mode == 0) {
    const uint4 gv = *reinterpret_cast<const uint4*>(second + img + ((size_t)yq * G + xq) * C + c0);
    float g[8];
    bw_unpack8(gv, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float u = acc[j];
      const float cdf = 0.5f * (1.0f + erff(u * 0.70710678118654752f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * u * u);
      acc[j] = g[j] * (cdf + u * pdf);
    }
  }
  uint4 o;
  o.x = bw_pk2(acc[0], acc[1]); o.y = bw_pk2(acc[2], acc[3]); o.z = bw_pk2(acc[4], acc[5]); o.w = bw_pk2(acc[6], acc[7]);
  *reinterpret_cast<uint4*>(out + img + ((size_t)yq * G + xq) * C + c0) = o;
}

// partial[chunk][10][C]: rows 0..8 = dw taps, row 9 = db; thread = channel pair, block = 128 channel pairs x 1 chunk
__global__ void __launch_bounds__(128) dwconv_bwd_dw_kernel(const bf16* __restrict__ h, const bf16* __restrict__ du,
                                                            float* __restrict__ partial, int B, int G, int C,
                                                            int pos_per_chunk) {
  const int c = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (c >= C) return;
  const long long total = (long long)B * G * G;
  const long long p0 = (long long)blockIdx.y * pos_per_chunk;
  const long long p1 = p0 + pos_per_chunk < total ? p0 + pos_per_chunk : total;
  float2 acc[10];
#pragma unroll
  for (int t = 0; t < 10; ++t) acc[t] = make_float2(0.f, 0.f);
  for (long long p = p0; p < p1; ++p) {
    const int xq = int(p % G), yq = int((p / G) % G);
    const long long b = p / ((long long)G * G);
    const size_t img = (size_t)b * G * G * C;
    const uint32_t dv = *reinterpret_cast<const uint32_t*>(du + img + ((size_t)yq * G + xq) * C + c);
    const float d0 = __uint_as_float(dv << 16), d1 = __uint_as_float(dv & 0xffff0000u);
    acc[9].x += d0; acc[9].y += d1;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int yy = yq + dy - 1, xx = xq + dx - 1;
        if (yy < 0 || yy >= G || xx < 0 || xx >= G) continue;
        const uint32_t hv = *reinterpret_cast<const uint32_t*>(h + img + ((size_t)yy * G + xx) * C + c);
        acc[dy * 3 + dx].x += d0 * __uint_as_float(hv << 16);
        acc[dy * 3 + dx].y += d1 * __uint_as_float(hv & 0xffff0000u);
      }
  }
#pragma unroll
  for (int t = 0; t < 10; ++t)
    *reinterpret_cast<float2*>(partial + ((size_t)blockIdx.y * 10 + t) * C + c) = acc[t];
}
__global__ void __launch_bounds__(256) dwconv_bwd_dw_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw9,
                                                                   float* __restrict__ db, int nchunk, int C) {
  const int i = blockIdx.x * 256 + threadIdx.x;  // over 10*C
  if (i >= 10 * C) return;
  float t = 0.f;
  for (int k = 0; k < nchunk; ++k) t += partial[(size_t)k * 10 * C + i];
  if (i < 9 * C) dw9[i] = t;
  else db[i - 9 * C] = t;
}


// hid [B,G,G,C] (pre-conv), dg [B,G,G,C] (grad of the GELU output) -> du_tmp (scratch, same shape), dhid, dw9 [9,C], db [C]
int launch_dwconv_gelu_bwd(const bf16* hid, const bf16* dg, const float* w9, const float* bias, bf16* du_tmp, bf16* dhid,
                           float* dw9, float* db, int B, int G, int C, cudaStream_t st) {
  TLD_CHECK(C % 8 == 0, "dwconv_bwd: channels must be a multiple of 8");
  const long long threads = (long long)B * G * G * (C / 8);
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  if (G == 16 && C % 64 == 0 && B <= 65535) {  // the 256-px model: shared-memory tile kernel (rowwise.cu)
    if (launch_dwconv_g16_bwd(hid, dg, w9, bias, du_tmp, B, C, 1, st)) return 1;
    if (launch_dwconv_g16_bwd(du_tmp, nullptr, w9, bias, dhid, B, C, 2, st)) return 1;
  } else {
    dwconv_bwd_kernel<<<blocks, 256, 0, st>>>(hid, dg, w9, bias, du_tmp, B, G, C, 0);
    TLD_CUDA_OK(cudaGetLastError());
    dwconv_bwd_kernel<<<blocks, 256, 0, st>>>(du_tmp, nullptr, w9, bias, dhid, B, G, C, 1);
    TLD_CUDA_OK(cudaGetLastError());
  }
  if (G == 16 && C % 64 == 0 && B <= 65535) {
    // tap / bias gradients with the tile kernel too: one [10][C] partial per image, then the fixed-order sum over images
    float* g_dw_partial = device_scratch(SCR_DWCONV_DW, (size_t)B * 10 * C);
    if (!g_dw_partial) return 1;
    if (launch_dwconv_g16_bwd(hid, du_tmp, w9, bias, reinterpret_cast<bf16*>(g_dw_partial), B, C, 3, st)) return 1;
    dwconv_bwd_dw_reduce_kernel<<<(10 * C + 255) / 256, 256, 0, st>>>(g_dw_partial, dw9, db, B, C);
    TLD_CUDA_OK(cudaGetLastError());
    return 0;
  }
  const long long total = (long long)B * G * G;
  int nchunk = (int)((total + 31) / 32);   // >= 32 positions per chunk, at most 256 chunks (12 x 256 CTAs at C = 3072)
  if (nchunk > 256) nchunk = 256;
  const int ppc = (int)((total + nchunk - 1) / nchunk);
  nchunk = (int)((total + ppc - 1) / ppc);
  float* g_dw_partial = device_scratch(SCR_DWCONV_DW, (size_t)nchunk * 10 * C);
  if (!g_dw_partial) return 1;
  dwconv_bwd_dw_kernel<<<dim3((C / 2 + 127) / 128, nchunk), 128, 0, st>>>(hid, du_tmp, g_dw_partial, B, G, C, ppc);
  TLD_CUDA_OK(cudaGetLastError());
  dwconv_bwd_dw_reduce_kernel<<<(10 * C + 255) / 256, 256, 0, st>>>(g_dw_partial, dw9, db, nchunk, C);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Cross-attention backward (transformer_blocks.py:62-72 with 2 cond tokens).  Per row r (sample b) and head h:
//   s_j = scale q.k_j, p = softmax(s0,s1), out = p0 v0 + p1 v1;   given go = d(out):
//   dp_j = go.v_j; ds_j = p_j (dp_j - (p0 dp0 + p1 dp1)); dq = scale (ds0 k0 + ds1 k1)
//   dk_j += scale ds_j q (over the sample's rows);  dv_j += p_j go
// Block = (head, sample, chunk of 256 rows); thread = row for the per-row part, then = (vector, dim) for the reductions.
// dkv0/dkv1 [B, 2D] (K | V) must be zero-initialised (atomicAdd across row chunks).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) xattn_bwd_kernel(const bf16* __restrict__ q, const float* __restrict__ go,
                                                        const float* __restrict__ kv0, const float* __restrict__ kv1,
                                                        long long kv_stride, bf16* __restrict__ dq,
                                                        float* __restrict__ dkv0, float* __restrict__ dkv1,
                                                        long long dkv_stride, int n_tok, int D) {
  __shared__ float s_kv[4][64];
  __shared__ float s_row[256][4];  // ds0*scale, ds1*scale, p0, p1
  const int head = blockIdx.x, b = blockIdx.y, chunk = blockIdx.z;
  const int tid = threadIdx.x;
  const float scale = 0.125f;
  {
    const int vec = tid >> 6, d = tid & 63;
    const float* src = (vec & 1) ? kv1 + (size_t)b * kv_stride : kv0 + (size_t)b * kv_stride;
    s_kv[vec][d] = src[(vec >= 2 ? D : 0) + head * 64 + d];  // 0:k0 1:k1 2:v0 3:v1
  }
  __syncthreads();
  const int rl = chunk * 256 + tid;
  const bool ok = rl < n_tok;
  const long long row = (long long)b * n_tok + rl;
  float ds0 = 0.f, ds1 = 0.f, p0 = 0.f, p1 = 0.f;
  if (ok) {
    const bf16* qr = q + row * D + head * 64;
    const float* gr = go + row * D + head * 64;
    float s0 = 0.f, s1 = 0.f, dp0 = 0.f, dp1 = 0.f;
    // the row's 64 q values (128 B) and 64 upstream gradients (256 B) as 16-byte vector loads
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const uint4 qv = *reinterpret_cast<const uint4*>(qr + v * 8);
      const float4 g0 = *reinterpret_cast<const float4*>(gr + v * 8), g1 = *reinterpret_cast<const float4*>(gr + v * 8 + 4);
      const uint32_t qw[4] = {qv.x, qv.y, qv.z, qv.w};
      const float gvv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int d = v * 8 + e;
        const float qf = (e & 1) ? __uint_as_float(qw[e >> 1] & 0xffff0000u) : __uint_as_float(qw[e >> 1] << 16);
        s0 += qf * s_kv[0][d]; s1 += qf * s_kv[1][d];
        dp0 += gvv[e] * s_kv[2][d]; dp1 += gvv[e] * s_kv[3][d];
      }
    }
    s0 *= scale; s1 *= scale;
    const float mx = fmaxf(s0, s1), e0 = __expf(s0 - mx), e1 = __expf(s1 - mx), inv = 1.f / (e0 + e1);
    p0 = e0 * inv; p1 = e1 * inv;
    const float dot = p0 * dp0 + p1 * dp1;
    ds0 = p0 * (dp0 - dot) * scale;
    ds1 = p1 * (dp1 - dot) * scale;
    bf16* dqr = dq + row * D + head * 64;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int d = v * 8 + 2 * e;
        o[e] = bw_pk2(ds0 * s_kv[0][d] + ds1 * s_kv[1][d], ds0 * s_kv[0][d + 1] + ds1 * s_kv[1][d + 1]);
      }
      *reinterpret_cast<uint4*>(dqr + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  s_row[tid][0] = ds0; s_row[tid][1] = ds1; s_row[tid][2] = p0; s_row[tid][3] = p1;
  __syncthreads();
  // reductions over the chunk's rows: thread = (vec, d): dk0, dk1 (weights ds_j, data q), dv0, dv1 (weights p_j, data go)
  const int vec = tid >> 6, d = tid & 63;
  const int rows_here = min(256, n_tok - chunk * 256);
  const long long base = ((long long)b * n_tok + chunk * 256) * D + head * 64 + d;
  float acc = 0.f;
  if (vec < 2) {
    for (int r = 0; r < rows_here; ++r) acc += s_row[r][vec] * __bfloat162float(q[base + (long long)r * D]);
  } else {
    for (int r = 0; r < rows_here; ++r) acc += s_row[r][vec] * go[base + (long long)r * D];
  }
  float* dst = ((vec & 1) ? dkv1 : dkv0) + (size_t)b * dkv_stride + (vec >= 2 ? D : 0) + head * 64 + d;
  atomicAdd(dst, acc);
}

int launch_xattn_bwd(const bf16* q, const float* go, const float* kv0, const float* kv1, long long kv_stride, bf16* dq,
                     float* dkv0, float* dkv1, long long dkv_stride, int B, int n_tok, int D, cudaStream_t st) {
  TLD_CHECK(D % 64 == 0 && B <= 65535, "xattn_bwd: bad shape");
  dim3 grid(D / 64, B, (n_tok + 255) / 256);
  xattn_bwd_kernel<<<grid, 256, 0, st>>>(q, go, kv0, kv1, kv_stride, dq, dkv0, dkv1, dkv_stride, n_tok, D);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace tld

// ---- C ABI (parity-test entry points for the backward ops) ----
extern "C" {
#define TLD_EXPORT __attribute__((visibility("default")))
TLD_EXPORT int tld_bwd_cast_transpose(const float* in, uint16_t* out, uint16_t* outT, int rows, int cols, void* stream) {
  return tld::launch_cast_transpose_f32(in, reinterpret_cast<tld::bf16*>(out), reinterpret_cast<tld::bf16*>(outT), rows, cols,
                                        reinterpret_cast<cudaStream_t>(stream));
}
TLD_EXPORT int tld_bwd_colsum(const float* in, float* out, int rows, int cols, void* stream) {
  return tld::launch_colsum_f32(in, out, rows, cols, 0, reinterpret_cast<cudaStream_t>(stream));
}
TLD_EXPORT int tld_bwd_layernorm(const float* dy, const float* x, const float* gamma, float* dx, float* dgamma, float* dbeta,
                                 int rows, int D, void* stream) {
  return tld::launch_layernorm_bwd(dy, x, gamma, dx, dgamma, dbeta, rows, D, reinterpret_cast<cudaStream_t>(stream));
}
TLD_EXPORT int tld_bwd_dwconv_gelu(const uint16_t* hid, const uint16_t* dg, const float* w9, const float* bias,
                                   uint16_t* du_tmp, uint16_t* dhid, float* dw9, float* db, int batch, int grid, int channels,
                                   void* stream) {
  return tld::launch_dwconv_gelu_bwd(reinterpret_cast<const tld::bf16*>(hid), reinterpret_cast<const tld::bf16*>(dg), w9, bias,
                                     reinterpret_cast<tld::bf16*>(du_tmp), reinterpret_cast<tld::bf16*>(dhid), dw9, db, batch,
                                     grid, channels, reinterpret_cast<cudaStream_t>(stream));
}
TLD_EXPORT int tld_bwd_xattn(const uint16_t* q, const float* go, const float* kv0, const float* kv1, uint16_t* dq, float* dkv0,
                             float* dkv1, int batch, int n_tok, int D, void* stream) {
  return tld::launch_xattn_bwd(reinterpret_cast<const tld::bf16*>(q), go, kv0, kv1, 2LL * D, reinterpret_cast<tld::bf16*>(dq),
                               dkv0, dkv1, 2LL * D, batch, n_tok, D, reinterpret_cast<cudaStream_t>(stream));
}
}
