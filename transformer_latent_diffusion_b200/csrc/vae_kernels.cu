// Row-wise kernels of the VAE decoder, NHWC bf16 activations:
//   GroupNorm(32 groups, eps) [+ SiLU] fused into two HBM passes (statistics, then normalise+affine+SiLU), and the
//   nearest-neighbour 2x upsample.  In the PyTorch library path these are 72 % of the decode time (ATen's
//   channels-last GroupNorm + separate SiLU/elementwise kernels); the 3x3 convolutions stay on cuDNN in round 1.
// Reference call site: tld/diffusion.py:91 -> diffusers AutoencoderKL.decode (ResnetBlock2D: GroupNorm -> SiLU ->
// conv3x3, UpSample2D: nearest 2x -> conv3x3).
#include <cuda_fp16.h>
#include "common.h"

namespace tld {

__device__ __forceinline__ void bf16x8_to_float(const uint4& v, float (&f)[8]) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pk2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// partial[b][blk][g] = (sum, sum of squares) of group g over the block's pixel range
__global__ void __launch_bounds__(256) gn_stats_kernel(const bf16* __restrict__ x, const float* __restrict__ pre_bias,
                                                       float2* __restrict__ partial, int HW, int C, int groups,
                                                       int pix_per_block) {
  __shared__ float s_sum[64], s_sq[64];
  const int b = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
  const int tpp = C >> 3;                 // threads per pixel (8 channels = 16 bytes each)
  const int cpg = C / groups;             // channels per group
  const int cv = threadIdx.x % tpp, poff = threadIdx.x / tpp, pstride = 256 / tpp;
  if (threadIdx.x < 64) { s_sum[threadIdx.x] = 0.f; s_sq[threadIdx.x] = 0.f; }
  __syncthreads();
  const int p0 = blk * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  const bf16* xb = x + (size_t)b * HW * C + cv * 8;
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;  // first / second half of the 8 channels
  float pb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) pb[j] = pre_bias ? pre_bias[cv * 8 + j] : 0.f;
  for (int p = p0 + poff; p < p1; p += pstride) {
    const uint4 v = *reinterpret_cast<const uint4*>(xb + (size_t)p * C);
    float f[8];
    bf16x8_to_float(v, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += pb[j];
    s0 += (f[0] + f[1]) + (f[2] + f[3]);
    q0 += (f[0] * f[0] + f[1] * f[1]) + (f[2] * f[2] + f[3] * f[3]);
    s1 += (f[4] + f[5]) + (f[6] + f[7]);
    q1 += (f[4] * f[4] + f[5] * f[5]) + (f[6] * f[6] + f[7] * f[7]);
  }
  const int g0 = (cv * 8) / cpg, g1 = (cv * 8 + 4) / cpg;
  if (g0 == g1) {
    atomicAdd(&s_sum[g0], s0 + s1);
    atomicAdd(&s_sq[g0], q0 + q1);
  } else {
    atomicAdd(&s_sum[g0], s0); atomicAdd(&s_sq[g0], q0);
    atomicAdd(&s_sum[g1], s1); atomicAdd(&s_sq[g1], q1);
  }
  __syncthreads();
  if (threadIdx.x < groups) partial[((size_t)b * nblk + blk) * groups + threadIdx.x] = make_float2(s_sum[threadIdx.x], s_sq[threadIdx.x]);
}

// y = act((x - mean_g) * rstd_g * gamma_c + beta_c), act = SiLU or identity
__global__ void __launch_bounds__(256) gn_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ pre_bias,
                                                       const float2* __restrict__ partial, int n_partial,
                                                       const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, bf16* __restrict__ y, int HW, int C,
                                                       int groups, float eps, int silu, int pix_per_block) {
  __shared__ float s_a[512], s_b[512], s_mean[64], s_rstd[64];
  const int b = blockIdx.y, blk = blockIdx.x;
  const int cpg = C / groups;
  if (threadIdx.x < groups) {
    double s = 0.0, q = 0.0;
    for (int i = 0; i < n_partial; ++i) {
      const float2 v = partial[((size_t)b * n_partial + i) * groups + threadIdx.x];
      s += v.x; q += v.y;
    }
    const double n = (double)HW * cpg;
    const double mean = s / n;
    const double var = fmax(q / n - mean * mean, 0.0);
    s_mean[threadIdx.x] = (float)mean;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    const int g = c / cpg;
    const float a = s_rstd[g] * gamma[c];
    s_a[c] = a;
    s_b[c] = beta[c] + ((pre_bias ? pre_bias[c] : 0.f) - s_mean[g]) * a;  // (x + pb - mean) * a + beta
  }
  __syncthreads();
  const int tpp = C >> 3;
  const int cv = threadIdx.x % tpp, poff = threadIdx.x / tpp, pstride = 256 / tpp;
  float a[8], bb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = s_a[cv * 8 + j]; bb[j] = s_b[cv * 8 + j]; }
  const int p0 = blk * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  const size_t base = (size_t)b * HW * C + cv * 8;
  for (int p = p0 + poff; p < p1; p += pstride) {
    const uint4 v = *reinterpret_cast<const uint4*>(x + base + (size_t)p * C);
    float f[8];
    bf16x8_to_float(v, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = fmaf(f[j], a[j], bb[j]);
      if (silu) t = t / (1.0f + __expf(-t));
      f[j] = t;
    }
    uint4 o;
    o.x = pk2(f[0], f[1]); o.y = pk2(f[2], f[3]); o.z = pk2(f[4], f[5]); o.w = pk2(f[6], f[7]);
    *reinterpret_cast<uint4*>(y + base + (size_t)p * C) = o;
  }
}

// nearest-neighbour 2x upsample, NHWC: each thread copies one 16-byte channel vector to its 4 output pixels
__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H,
                                                         int W, int C8) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * W * C8;
  if (idx >= total) return;
  const int c = int(idx % C8);
  const long long pix = idx / C8;
  const int xw = int(pix % W);
  const int yh = int((pix / W) % H);
  const long long b = pix / ((long long)W * H);
  const uint4 v = x[idx];
  const long long orow = ((b * 2 * H + 2 * yh) * 2 * W + 2 * xw) * C8 + c;
  y[orow] = v;
  y[orow + C8] = v;
  y[orow + 2LL * W * C8] = v;
  y[orow + 2LL * W * C8 + C8] = v;
}

// out = x + h + bias[c]  (ResnetBlock2D tail: shortcut + conv2 output, with conv2's bias folded in)
__global__ void __launch_bounds__(256) add_bias_kernel(const uint4* __restrict__ x, const uint4* __restrict__ h,
                                                       const float* __restrict__ bias, uint4* __restrict__ out,
                                                       long long n_vec, int C8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vec) return;
  const int c = int(i % C8) * 8;
  float a[8], b[8];
  bf16x8_to_float(x[i], a);
  bf16x8_to_float(h[i], b);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] += b[j] + (bias ? bias[c + j] : 0.f);
  uint4 o;
  o.x = pk2(a[0], a[1]); o.y = pk2(a[2], a[3]); o.z = pk2(a[4], a[5]); o.w = pk2(a[6], a[7]);
  out[i] = o;
}

// ------------------------------------------------------------------------------------------------------------------
// decoder.conv_out: 3x3 'same' convolution 128 -> 3 channels at full resolution (diffusers AutoencoderKL.decode, called at
// tld/diffusion.py:91).  With 3 output channels there is nothing for a tensor core to do (an implicit GEMM padded to N = 64
// re-loads the 268 MB input once per tap: 1.96 ms per 16 images); this is an HBM-bound direct convolution on the CUDA cores:
// one CTA = 8 x 32 output pixels, thread = pixel, the haloed 10 x 34 input tile goes through shared memory 32 channels at a
// time TRANSPOSED to [channel pair][pixel] (conflict-free LDS.32 for lanes = consecutive pixels), the 3456 weights sit in the
// kernel-parameter constant bank, so every FMA takes its weight as a constant operand (no weight loads in the loop).
// Per 16 images of 256 x 256: 268 MB read once (x 1.33 halo), 3.6 GFMA -> ~0.15 ms.  Output: fp32 NCHW, the final image.
// ------------------------------------------------------------------------------------------------------------------
constexpr int TO_CIN = 128, TO_COUT = 3, TO_TH = 8, TO_TW = 32, TO_HW = (TO_TH + 2) * (TO_TW + 2);
struct ThinConvW {
  float w[9 * TO_CIN * TO_COUT];   // [tap][cin][cout]
  float b[4];
};
__global__ void __launch_bounds__(256) conv3x3_thin_out_kernel(const bf16* __restrict__ x, float* __restrict__ out,
                                                               const __grid_constant__ ThinConvW wt, int H, int W) {
  __shared__ uint32_t s_x[16][TO_HW + 1];   // [channel pair of the 32-channel chunk][haloed pixel]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x0 = blockIdx.x * TO_TW, y0 = blockIdx.y * TO_TH, b = blockIdx.z;
  const bf16* xb = x + (size_t)b * H * W * TO_CIN;
  float acc[TO_COUT] = {wt.b[0], wt.b[1], wt.b[2]};
#pragma unroll
  for (int chunk = 0; chunk < TO_CIN / 32; ++chunk) {
    __syncthreads();   // the previous chunk has been consumed
    // 340 haloed pixels x 4 pieces of 16 bytes (8 channels): piece-major so that consecutive threads take consecutive pixels
    for (int i = threadIdx.x; i < TO_HW * 4; i += 256) {
      const int piece = i / TO_HW, p = i - piece * TO_HW;
      const int py = p / (TO_TW + 2), px = p - py * (TO_TW + 2);
      const int gy = y0 + py - 1, gx = x0 + px - 1;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = *reinterpret_cast<const uint4*>(xb + ((size_t)gy * W + gx) * TO_CIN + chunk * 32 + piece * 8);
      s_x[piece * 4 + 0][p] = v.x;
      s_x[piece * 4 + 1][p] = v.y;
      s_x[piece * 4 + 2][p] = v.z;
      s_x[piece * 4 + 3][p] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int cp = 0; cp < 16; ++cp)
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const uint32_t v = s_x[cp][(ty + tap / 3) * (TO_TW + 2) + tx + tap % 3];
        const float a0 = __uint_as_float(v << 16), a1 = __uint_as_float(v & 0xffff0000u);
        const int c = chunk * 32 + cp * 2;
#pragma unroll
        for (int o = 0; o < TO_COUT; ++o) {
          acc[o] = fmaf(a0, wt.w[(tap * TO_CIN + c) * TO_COUT + o], acc[o]);
          acc[o] = fmaf(a1, wt.w[(tap * TO_CIN + c + 1) * TO_COUT + o], acc[o]);
        }
      }
  }
  const int gy = y0 + ty, gx = x0 + tx;
  if (gy < H && gx < W) {
#pragma unroll
    for (int o = 0; o < TO_COUT; ++o) out[(((size_t)b * TO_COUT + o) * H + gy) * W + gx] = acc[o];
  }
}

// x NHWC bf16 [B,H,W,128] (device); weights [3,128,3,3] (OIHW) + bias [3] as HOST fp32 (they become kernel parameters)
int launch_conv3x3_thin_out(const bf16* x, const float* w_host_oihw, const float* b_host, float* out, int B, int H, int W,
                            cudaStream_t st) {
  TLD_CHECK(B > 0 && B <= 65535 && H > 0 && W > 0, "conv3x3_thin_out: bad shape");
  ThinConvW wt;
  for (int o = 0; o < TO_COUT; ++o)
    for (int c = 0; c < TO_CIN; ++c)
      for (int t = 0; t < 9; ++t) wt.w[(t * TO_CIN + c) * TO_COUT + o] = w_host_oihw[(o * TO_CIN + c) * 9 + t];
  wt.b[0] = b_host[0]; wt.b[1] = b_host[1]; wt.b[2] = b_host[2]; wt.b[3] = 0.f;
  dim3 grid((W + TO_TW - 1) / TO_TW, (H + TO_TH - 1) / TO_TH, B);
  conv3x3_thin_out_kernel<<<grid, 256, 0, st>>>(x, out, wt, H, W);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_add_bias(const bf16* x, const bf16* h, const float* bias, bf16* out, long long n, int C, cudaStream_t st) {
  TLD_CHECK(C % 8 == 0 && n % 8 == 0, "add_bias: channels must be a multiple of 8");
  const long long nv = n / 8;
  add_bias_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(x),
                                                                reinterpret_cast<const uint4*>(h), bias,
                                                                reinterpret_cast<uint4*>(out), nv, C / 8);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_gn_act(const bf16* x, const float* pre_bias, const float* gamma, const float* beta, bf16* y, int B, int HW,
                  int C, int groups, float eps, int silu, cudaStream_t st) {
  TLD_CHECK(C % 8 == 0 && C <= 512 && (256 % (C / 8)) == 0, "group_norm: channels must be 128/256/512-like (C/8 divides 256)");
  TLD_CHECK(groups <= 64 && C % groups == 0 && (C / groups) % 4 == 0, "group_norm: bad group configuration");
  TLD_CHECK(B <= 65535, "group_norm: batch too large");
  // enough blocks to fill the machine, at least 256 pixels per block
  int nblk = (HW + 2047) / 2048;
  const int want = (4 * sm_count() + B - 1) / B;
  if (nblk < want) nblk = want;
  int ppb = (HW + nblk - 1) / nblk;
  const int pstride = 256 / (C / 8);
  ppb = ((ppb + pstride - 1) / pstride) * pstride;
  nblk = (HW + ppb - 1) / ppb;
  const size_t need = (size_t)B * nblk * groups;
  float2* g_partial = reinterpret_cast<float2*>(device_scratch(SCR_GROUPNORM, 2 * need));
  if (!g_partial) return 1;
  gn_stats_kernel<<<dim3(nblk, B), 256, 0, st>>>(x, pre_bias, g_partial, HW, C, groups, ppb);
  TLD_CUDA_OK(cudaGetLastError());
  gn_apply_kernel<<<dim3(nblk, B), 256, 0, st>>>(x, pre_bias, g_partial, nblk, gamma, beta, y, HW, C, groups, eps, silu, ppb);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

// GroupNorm whose statistics partials came out of the producing convolution's epilogue (gemm_tcgen05.cuh, EPI_BIAS_BF16 with
// gn_part): qpart [B * HW / 32 slabs][C / 4 quads] (sum, sum of squares) -> one (sum, sum of squares) per (image, group),
// then the ordinary apply pass.  The tensor is read once instead of twice; the reduction order is fixed (deterministic).
__global__ void __launch_bounds__(256) gn_finalize_kernel(const float2* __restrict__ qpart, float2* __restrict__ out,
                                                          int slabs_per_image, int Q, int groups) {
  __shared__ double s_s[256], s_q[256];
  const int b = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;   // block blk sums slabs [blk * per, (blk + 1) * per)
  const int per = (slabs_per_image + nblk - 1) / nblk;
  const int i0 = blk * per, i1 = min(slabs_per_image, i0 + per);
  const int quad = threadIdx.x % Q, sub = threadIdx.x / Q, nsub = 256 / Q;
  double s = 0.0, q = 0.0;
  const float2* p = qpart + (size_t)b * slabs_per_image * Q + quad;
  for (int i = i0 + sub; i < i1; i += nsub) {
    const float2 v = p[(size_t)i * Q];
    s += v.x;
    q += v.y;
  }
  s_s[threadIdx.x] = s;
  s_q[threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.x < groups) {   // group g = quads [g * qpg, (g + 1) * qpg) x all sub-lanes, summed in a fixed order
    const int qpg = Q / groups;
    double ts = 0.0, tq = 0.0;
    for (int u = 0; u < nsub; ++u)
      for (int k = 0; k < qpg; ++k) {
        ts += s_s[u * Q + threadIdx.x * qpg + k];
        tq += s_q[u * Q + threadIdx.x * qpg + k];
      }
    out[((size_t)b * nblk + blk) * groups + threadIdx.x] = make_float2((float)ts, (float)tq);
  }
}

int launch_gn_act_from_partials(const bf16* x, const float* qpart, const float* gamma, const float* beta, bf16* y, int B, int HW,
                                int C, int groups, float eps, int silu, cudaStream_t st) {
  TLD_CHECK(C % 8 == 0 && C <= 512 && (256 % (C / 8)) == 0, "group_norm: channels must be 128/256/512-like (C/8 divides 256)");
  TLD_CHECK(groups <= 64 && C % groups == 0 && (C / groups) % 4 == 0, "group_norm: bad group configuration");
  TLD_CHECK(B <= 65535 && HW % 32 == 0 && qpart != nullptr, "group_norm: conv-epilogue partials need H*W % 32 == 0");
  const int Q = C / 4;
  TLD_CHECK(Q <= 256 && 256 % Q == 0 && Q % groups == 0, "group_norm: bad quad layout");
  // split every image's slabs over enough blocks to fill the machine (the apply kernel sums the nsplit partials per group)
  int nsplit = (2 * sm_count() + B - 1) / B;
  if (nsplit > HW / 32 / 8) nsplit = HW / 32 / 8;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > 64) nsplit = 64;
  float2* stats = reinterpret_cast<float2*>(device_scratch(SCR_GROUPNORM, 2 * (size_t)B * nsplit * groups));
  if (!stats) return 1;
  gn_finalize_kernel<<<dim3(nsplit, B), 256, 0, st>>>(reinterpret_cast<const float2*>(qpart), stats, HW / 32, Q, groups);
  TLD_CUDA_OK(cudaGetLastError());
  int nblk = (HW + 2047) / 2048;
  const int want = (4 * sm_count() + B - 1) / B;
  if (nblk < want) nblk = want;
  int ppb = (HW + nblk - 1) / nblk;
  const int pstride = 256 / (C / 8);
  ppb = ((ppb + pstride - 1) / pstride) * pstride;
  nblk = (HW + ppb - 1) / ppb;
  gn_apply_kernel<<<dim3(nblk, B), 256, 0, st>>>(x, nullptr, stats, nsplit, gamma, beta, y, HW, C, groups, eps, silu, ppb);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

// Mid-block attention core (one head as wide as the channel count, diffusers Attention with heads = 1): per image
// S = Q K^T (tcgen05 GEMM, fp32 out) -> P = softmax(S / sqrt(C)) (row kernel, bf16) -> O = P V (tcgen05 GEMM with V as the
// [K, N]-stored operand).  1024 tokens x 512 channels per image: 3 small launches per image, ~1 ms per 64-image decode.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, bf16* __restrict__ p, int rows, int cols,
                                                           float scale_log2e) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* sr = s + (size_t)row * cols;
  float mx = -INFINITY;
  for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, sr[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int c = lane; c < cols; c += 32) sum += exp2f((sr[c] - mx) * scale_log2e);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  bf16* pr = p + (size_t)row * cols;
  for (int c = lane; c < cols; c += 32) pr[c] = __float2bfloat16(exp2f((sr[c] - mx) * scale_log2e) * inv);
}

int launch_vae_attention_core(const bf16* q, const bf16* k, const bf16* v, bf16* out, int B, int n, int C, cudaStream_t st) {
  TLD_CHECK(q && k && v && out && B > 0, "vae_attention: null argument");
  TLD_CHECK(n % 64 == 0 && C % 64 == 0, "vae_attention: tokens and channels must be multiples of 64");
  float* scr = device_scratch(SCR_VAE_ATTN, (size_t)n * n + (size_t)n * n / 2);
  if (!scr) return 1;
  float* s = scr;
  bf16* p = reinterpret_cast<bf16*>(scr + (size_t)n * n);
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)C);
  for (int b = 0; b < B; ++b) {
    const size_t off = (size_t)b * n * C;
    if (launch_gemm(4 /* EPI_F32 */, q + off, C, k + off, C, n, n, C, s, n, nullptr, nullptr, st)) return 1;
    softmax_rows_kernel<<<(n + 7) / 8, 256, 0, st>>>(s, p, n, n, scale_log2e);
    TLD_CUDA_OK(cudaGetLastError());
    if (launch_gemm_nn(0 /* EPI_BF16 */, p, n, v + off, C, n, C, n, out + off, C, st)) return 1;
  }
  return 0;
}

int launch_upsample2x(const bf16* x, bf16* y, int B, int H, int W, int C, cudaStream_t st) {
  TLD_CHECK(C % 8 == 0, "upsample: channels must be a multiple of 8");
  const long long total = (long long)B * H * W * (C / 8);
  upsample2x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(x),
                                                                    reinterpret_cast<uint4*>(y), B, H, W, C / 8);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Image post-processing on the device (tld/diffusion.py:185 `to_pil(make_grid((out + 1) / 2, nrow, padding=4).clip(0, 1))`,
// tld/train.py:36): decoded images [B,3,H,W] in [-1,1] -> ONE uint8 HWC grid (torchvision.make_grid layout: `ncol` per
// row, `pad` black pixels around every image), value = trunc(clip((x + 1) / 2, 0, 1) * 255) exactly as
// ToPILImage's mul(255).byte().  The host then copies 1 byte per sample instead of 4.
// For bf16 inputs the (x + 1) / 2 is rounded to bf16 after each operation, as torch does on a bf16 tensor.
// ------------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void __launch_bounds__(256) image_grid_u8_kernel(const TIn* __restrict__ img, uint8_t* __restrict__ out, int B, int H,
                                                            int W, int ncol, int pad, int GH, int GW) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;   // over GH * GW pixels
  if (i >= (long long)GH * GW) return;
  const int gy = int(i / GW), gx = int(i % GW);
  const int cell_h = H + pad, cell_w = W + pad;
  const int r = (gy - pad) / cell_h, c = (gx - pad) / cell_w;
  const int y = gy - pad - r * cell_h, x = gx - pad - c * cell_w;
  const int b = r * ncol + c;
  uint8_t v[3] = {0, 0, 0};
  if (gy >= pad && gx >= pad && y < H && x < W && c < ncol && b < B) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float f;
      if constexpr (sizeof(TIn) == 2) {
        const bf16 t = __float2bfloat16(__bfloat162float(img[(((size_t)b * 3 + ch) * H + y) * W + x]) + 1.0f);
        f = __bfloat162float(__float2bfloat16(__bfloat162float(t) * 0.5f));
      } else {
        f = (img[(((size_t)b * 3 + ch) * H + y) * W + x] + 1.0f) * 0.5f;
      }
      f = fminf(fmaxf(f, 0.f), 1.f);
      v[ch] = (uint8_t)(f * 255.0f);
    }
  }
  out[i * 3 + 0] = v[0];
  out[i * 3 + 1] = v[1];
  out[i * 3 + 2] = v[2];
}

// ------------------------------------------------------------------------------------------------
// uint8 latent storage of the dataset files (tld/data.py:51-60), HBM-bound byte work: 4 latents per thread.
//   quantize:   q = trunc(((clip(x, -c, c) / c + 1) / 2) * 255)      fp32 input: fp32 arithmetic (IEEE, no contraction);
//                                                                   fp16 input: every step rounded to fp16, as torch does
//   dequantize: x = ((half(q) / 255) * 2 - 1) * c                    fp16 out, every step rounded to fp16
// Bit-exact against vectors produced by the reference's own source (tests/golden/latent_quant.npz).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float h_round(float v) { return __half2float(__float2half_rn(v)); }

template <bool HALF_IN>
__global__ void __launch_bounds__(256) latent_quantize_kernel(const void* __restrict__ in, uint8_t* __restrict__ out, long long n,
                                                              float c) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float v;
  if constexpr (HALF_IN) {
    v = __half2float(reinterpret_cast<const __half*>(in)[i]);
    v = fminf(fmaxf(v, -c), c);
    v = h_round(__fdiv_rn(v, c));
    v = h_round(__fadd_rn(v, 1.0f));
    v = h_round(__fmul_rn(v, 0.5f));
    v = h_round(__fmul_rn(v, 255.0f));
  } else {
    v = reinterpret_cast<const float*>(in)[i];
    v = fminf(fmaxf(v, -c), c);
    v = __fdiv_rn(v, c);
    v = __fadd_rn(v, 1.0f);
    v = __fmul_rn(v, 0.5f);
    v = __fmul_rn(v, 255.0f);
  }
  out[i] = (uint8_t)v;
}

__global__ void __launch_bounds__(256) latent_dequantize_kernel(const uint8_t* __restrict__ q, __half* __restrict__ out,
                                                                long long n, float c) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float v = (float)q[i];
  v = h_round(__fdiv_rn(v, 255.0f));
  v = h_round(__fmul_rn(v, 2.0f));
  v = h_round(__fadd_rn(v, -1.0f));
  out[i] = __float2half_rn(__fmul_rn(v, c));
}

int launch_latent_quantize(const void* in, int is_fp16, uint8_t* out, long long n, float clip_val, cudaStream_t st) {
  TLD_CHECK(n >= 0 && clip_val > 0.f, "latent_quantize: bad arguments");
  if (n == 0) return 0;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (is_fp16) latent_quantize_kernel<true><<<blocks, 256, 0, st>>>(in, out, n, clip_val);
  else latent_quantize_kernel<false><<<blocks, 256, 0, st>>>(in, out, n, clip_val);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}
int launch_latent_dequantize(const uint8_t* q, void* out_fp16, long long n, float clip_val, cudaStream_t st) {
  TLD_CHECK(n >= 0 && clip_val > 0.f, "latent_dequantize: bad arguments");
  if (n == 0) return 0;
  latent_dequantize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(q, reinterpret_cast<__half*>(out_fp16), n, clip_val);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_image_grid_u8(const void* img, int is_bf16, uint8_t* out, int B, int H, int W, int ncol, int pad, cudaStream_t st) {
  TLD_CHECK(B > 0 && H > 0 && W > 0 && ncol > 0 && pad >= 0, "image_grid_u8: bad shape");
  const int nc = ncol < B ? ncol : B, nr = (B + nc - 1) / nc;
  const int GH = nr * (H + pad) + pad, GW = nc * (W + pad) + pad;
  const long long px = (long long)GH * GW;
  const unsigned blocks = (unsigned)((px + 255) / 256);
  if (is_bf16)
    image_grid_u8_kernel<bf16><<<blocks, 256, 0, st>>>(reinterpret_cast<const bf16*>(img), out, B, H, W, nc, pad, GH, GW);
  else
    image_grid_u8_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(img), out, B, H, W, nc, pad, GH, GW);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace tld

extern "C" {

// decoder.conv_out (128 -> 3, 3x3 'same'): x NHWC bf16 device, w_host [3,128,3,3] / b_host [3] HOST fp32 (they travel in the
// kernel-parameter constant bank), out fp32 NCHW device [batch,3,h,w]
__attribute__((visibility("default"))) int tld_vae_conv_out3(const uint16_t* x, const float* w_host, const float* b_host, float* out, int batch, int h, int w,
                              void* stream) {
  TLD_CHECK(x && w_host && b_host && out, "tld_vae_conv_out3: null argument");
  return tld::launch_conv3x3_thin_out(reinterpret_cast<const tld::bf16*>(x), w_host, b_host, out, batch, h, w,
                                      reinterpret_cast<cudaStream_t>(stream));
}

__attribute__((visibility("default"))) int tld_vae_group_norm(const uint16_t* x, const float* pre_bias,
                                                              const float* gamma, const float* beta, uint16_t* y,
                                                              int batch, int hw, int channels, int groups, float eps,
                                                              int silu, void* stream) {
  return tld::launch_gn_act(reinterpret_cast<const tld::bf16*>(x), pre_bias, gamma, beta, reinterpret_cast<tld::bf16*>(y),
                            batch, hw, channels, groups, eps, silu, reinterpret_cast<cudaStream_t>(stream));
}
__attribute__((visibility("default"))) int tld_vae_group_norm_from_conv(const uint16_t* x, const float* conv_partials,
                                                                        const float* gamma, const float* beta, uint16_t* y,
                                                                        int batch, int hw, int channels, int groups,
                                                                        float eps, int silu, void* stream) {
  return tld::launch_gn_act_from_partials(reinterpret_cast<const tld::bf16*>(x), conv_partials, gamma, beta,
                                          reinterpret_cast<tld::bf16*>(y), batch, hw, channels, groups, eps, silu,
                                          reinterpret_cast<cudaStream_t>(stream));
}
__attribute__((visibility("default"))) int tld_vae_conv3x3_fused(const uint16_t* x, const uint16_t* w, const float* bias,
                                                                 uint16_t* out, int batch, int h, int w_px, int cin, int cout,
                                                                 const uint16_t* residual, float* gn_partials, void* stream) {
  return tld::launch_conv3x3(reinterpret_cast<const tld::bf16*>(x), reinterpret_cast<const tld::bf16*>(w), bias,
                             reinterpret_cast<tld::bf16*>(out), batch, h, w_px, cin, cout,
                             reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const tld::bf16*>(residual), gn_partials);
}
__attribute__((visibility("default"))) int tld_vae_attention_core(const uint16_t* q, const uint16_t* k, const uint16_t* v,
                                                                  uint16_t* out, int batch, int n_tok, int channels,
                                                                  void* stream) {
  return tld::launch_vae_attention_core(reinterpret_cast<const tld::bf16*>(q), reinterpret_cast<const tld::bf16*>(k),
                                        reinterpret_cast<const tld::bf16*>(v), reinterpret_cast<tld::bf16*>(out), batch, n_tok,
                                        channels, reinterpret_cast<cudaStream_t>(stream));
}
__attribute__((visibility("default"))) int tld_vae_add_bias(const uint16_t* x, const uint16_t* h, const float* bias,
                                                            uint16_t* out, long long numel, int channels, void* stream) {
  return tld::launch_add_bias(reinterpret_cast<const tld::bf16*>(x), reinterpret_cast<const tld::bf16*>(h), bias,
                              reinterpret_cast<tld::bf16*>(out), numel, channels, reinterpret_cast<cudaStream_t>(stream));
}
__attribute__((visibility("default"))) int tld_vae_conv3x3(const uint16_t* x, const uint16_t* w, const float* bias,
                                                           uint16_t* out, int batch, int h, int w_px, int cin, int cout,
                                                           void* stream) {
  return tld::launch_conv3x3(reinterpret_cast<const tld::bf16*>(x), reinterpret_cast<const tld::bf16*>(w), bias,
                             reinterpret_cast<tld::bf16*>(out), batch, h, w_px, cin, cout,
                             reinterpret_cast<cudaStream_t>(stream));
}
__attribute__((visibility("default"))) int tld_latent_quantize(const void* lat, int is_fp16, uint8_t* out, long long n,
                                                               float clip_val, void* stream) {
  return tld::launch_latent_quantize(lat, is_fp16, out, n, clip_val, reinterpret_cast<cudaStream_t>(stream));
}
__attribute__((visibility("default"))) int tld_latent_dequantize(const uint8_t* q, uint16_t* out_fp16, long long n,
                                                                 float clip_val, void* stream) {
  return tld::launch_latent_dequantize(q, out_fp16, n, clip_val, reinterpret_cast<cudaStream_t>(stream));
}
__attribute__((visibility("default"))) int tld_image_grid_u8(const void* img, int is_bf16, uint8_t* out, int batch, int h,
                                                             int w, int ncol, int pad, void* stream) {
  return tld::launch_image_grid_u8(img, is_bf16, out, batch, h, w, ncol, pad, reinterpret_cast<cudaStream_t>(stream));
}
__attribute__((visibility("default"))) int tld_vae_upsample2x(const uint16_t* x, uint16_t* y, int batch, int h, int w,
                                                              int channels, void* stream) {
  return tld::launch_upsample2x(reinterpret_cast<const tld::bf16*>(x), reinterpret_cast<tld::bf16*>(y), batch, h, w,
                                channels, reinterpret_cast<cudaStream_t>(stream));
}
}
