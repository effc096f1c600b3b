// Self-attention backward (head_dim 64, non-causal), reference: autograd through F.scaled_dot_product_attention at
// tld/transformer_blocks.py:37-44.  Forward: O = softmax(Q K^T / 8) V, added straight into the residual stream, so
// dO is the residual gradient and O = x_after - x_before.
//
// Up to 256 tokens per sample (the 256-px model): one CTA = one (sample, head) with all keys resident in shared memory; more
// tokens: the key-tiled kernels further down.
// 8 warps, warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate):
//   phase 0  load Q,K,V (bf16) and dO (fp32 -> bf16) tiles, delta[r] = sum_d dO[r,d] * O[r,d]
//   phase 1  row log-sum-exp: warp w owns query rows [32w, 32w+32): S = Q K^T over all keys, online max/sum
//   phase 2  warp w owns keys [32w, 32w+32); for every block of 32 query rows:
//              S^T = K_w Q_b^T, dP^T = V_w dO_b^T           (keys x rows, so P^T / dS^T come out as A operands)
//              P^T = exp2(S^T c - lse[r]), dS^T = P^T (dP^T - delta[r])
//              dV_w += P^T dO_b,  dK_w += dS^T Q_b            (register accumulators for the whole kernel)
//              dS^T -> smem, then dQ_b = dS K with the 8 warps splitting the 32 x 64 output tile
// Outputs dq|dk|dv are written as bf16 into a [T, 3D] buffer laid out like qkv (the A operand of the qkv wgrad/dgrad).
#include "common.h"

namespace tld {

constexpr int AB_HD = 64;
constexpr int AB_MAXN = 256;
constexpr int AB_THREADS = 256;

__device__ __forceinline__ void ab_ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ab_ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ab_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t ab_pack(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// [rows][64] bf16 tile, 128 B per row, 16-byte chunk c of row r stored at chunk (c ^ (r & 7))
__device__ __forceinline__ bf16* ab_tile(bf16* base, int r, int chunk) { return base + r * AB_HD + ((chunk ^ (r & 7)) << 3); }
// dS^T tile [keys][32 rows] bf16, 64 B per row (4 chunks), chunk c of key k stored at chunk (c ^ ((k >> 1) & 3))
__device__ __forceinline__ bf16* ab_ds(bf16* base, int key, int chunk) { return base + key * 32 + ((chunk ^ ((key >> 1) & 3)) << 3); }

__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ d_out, const float* __restrict__ x_before,
                     const float* __restrict__ x_after, bf16* __restrict__ dqkv, int n_tok, int D) {
  extern __shared__ __align__(128) uint8_t ab_smem[];
  bf16* sQ = reinterpret_cast<bf16*>(ab_smem);
  bf16* sK = sQ + AB_MAXN * AB_HD;
  bf16* sV = sK + AB_MAXN * AB_HD;
  bf16* sDO = sV + AB_MAXN * AB_HD;
  bf16* sDS = sDO + AB_MAXN * AB_HD;                       // [256 keys][32 rows]
  float* s_lse = reinterpret_cast<float*>(sDS + AB_MAXN * 32);  // log2-domain: m*c + log2(l)
  float* s_delta = s_lse + AB_MAXN;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int head = blockIdx.x, b = blockIdx.y;
  const long long ld = 3LL * D;
  const long long row0 = (long long)b * n_tok;
  const float sl2 = 0.125f * 1.4426950408889634f;

  // ---- phase 0: tiles into smem ----
  for (int i = tid; i < n_tok * 8; i += AB_THREADS) {
    const int r = i >> 3, c = i & 7;
    const bf16* src = qkv + (row0 + r) * ld + head * AB_HD + c * 8;
    *reinterpret_cast<uint4*>(ab_tile(sQ, r, c)) = *reinterpret_cast<const uint4*>(src);
    *reinterpret_cast<uint4*>(ab_tile(sK, r, c)) = *reinterpret_cast<const uint4*>(src + D);
    *reinterpret_cast<uint4*>(ab_tile(sV, r, c)) = *reinterpret_cast<const uint4*>(src + 2 * D);
    const float* go = d_out + (row0 + r) * D + head * AB_HD + c * 8;
    const float* xa = x_after + (row0 + r) * D + head * AB_HD + c * 8;
    const float* xb = x_before + (row0 + r) * D + head * AB_HD + c * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(go), g1 = *reinterpret_cast<const float4*>(go + 4);
    const float4 a0 = *reinterpret_cast<const float4*>(xa), a1 = *reinterpret_cast<const float4*>(xa + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(xb), b1 = *reinterpret_cast<const float4*>(xb + 4);
    uint4 pk;
    pk.x = ab_pack(g0.x, g0.y); pk.y = ab_pack(g0.z, g0.w); pk.z = ab_pack(g1.x, g1.y); pk.w = ab_pack(g1.z, g1.w);
    *reinterpret_cast<uint4*>(ab_tile(sDO, r, c)) = pk;
    float part = g0.x * (a0.x - b0.x) + g0.y * (a0.y - b0.y) + g0.z * (a0.z - b0.z) + g0.w * (a0.w - b0.w) +
                 g1.x * (a1.x - b1.x) + g1.y * (a1.y - b1.y) + g1.z * (a1.z - b1.z) + g1.w * (a1.w - b1.w);
    // the 8 chunk-threads of a row are 8 consecutive lanes
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    if (c == 0) s_delta[r] = part;
  }
  __syncthreads();

  // ---- phase 1: lse per query row (warp owns 32 rows = 2 m-tiles) ----
  for (int mt = 0; mt < 2; ++mt) {
    const int qrow = warp * 32 + mt * 16;
    if (qrow >= n_tok) break;
    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ab_ldsm_x4(qf[ks], ab_tile(sQ, qrow + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4)));
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    for (int kc = 0; kc < n_tok; kc += 64) {
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t kf[4];
          ab_ldsm_x4(kf, ab_tile(sK, kc + np * 16 + (lane & 7) + (lane >> 4) * 8, ks * 2 + ((lane >> 3) & 1)));
          ab_mma(s[2 * np], qf[ks], kf[0], kf[1]);
          ab_mma(s[2 * np + 1], qf[ks], kf[2], kf[3]);
        }
      float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        mx[0] = fmaxf(mx[0], fmaxf(s[i][0], s[i][1]));
        mx[1] = fmaxf(mx[1], fmaxf(s[i][2], s[i][3]));
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      }
      float rs[2] = {0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        rs[0] += exp2f((s[i][0] - mx[0]) * sl2) + exp2f((s[i][1] - mx[0]) * sl2);
        rs[1] += exp2f((s[i][2] - mx[1]) * sl2) + exp2f((s[i][3] - mx[1]) * sl2);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        l_run[r] = l_run[r] * exp2f((m_run[r] - mx[r]) * sl2) + rs[r];
        m_run[r] = mx[r];
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
      l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    if (t4 == 0) {
      s_lse[qrow + g] = m_run[0] * sl2 + log2f(l_run[0]);
      s_lse[qrow + g + 8] = m_run[1] * sl2 + log2f(l_run[1]);
    }
  }
  __syncthreads();

  // ---- phase 2 ----
  const int key0 = warp * 32;              // this warp's key slice
  const bool kv_active = key0 < n_tok;
  float dV[2][8][4], dK[2][8][4];          // [key m-tile][d n-tile][frag]
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dV[a][i][j] = dK[a][i][j] = 0.f;

  for (int qb = 0; qb < n_tok; qb += 32) {
    if (kv_active) {
      // S^T = K_w Q_b^T and dP^T = V_w dO_b^T : M = 32 keys (2 m-tiles), N = 32 rows (4 n-tiles), K = 64
      float st[2][4][4], dpt[2][4][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) st[a][i][j] = dpt[a][i][j] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t kf[2][4], vf[2][4];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int r = key0 + a * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          ab_ldsm_x4(kf[a], ab_tile(sK, r, ks * 2 + (lane >> 4)));
          ab_ldsm_x4(vf[a], ab_tile(sV, r, ks * 2 + (lane >> 4)));
        }
#pragma unroll
        for (int np = 0; np < 2; ++np) {  // pairs of 8-row n-tiles
          uint32_t qf[4], gf[4];
          const int r = qb + np * 16 + (lane & 7) + (lane >> 4) * 8;
          ab_ldsm_x4(qf, ab_tile(sQ, r, ks * 2 + ((lane >> 3) & 1)));
          ab_ldsm_x4(gf, ab_tile(sDO, r, ks * 2 + ((lane >> 3) & 1)));
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            ab_mma(st[a][2 * np], kf[a], qf[0], qf[1]);
            ab_mma(st[a][2 * np + 1], kf[a], qf[2], qf[3]);
            ab_mma(dpt[a][2 * np], vf[a], gf[0], gf[1]);
            ab_mma(dpt[a][2 * np + 1], vf[a], gf[2], gf[3]);
          }
        }
      }
      // P^T and dS^T as A-operand fragments: k index = query rows (2 k-steps of 16 rows)
      uint32_t pf[2][2][4], dsf[2][2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int rr = qb + nt * 8 + t4 * 2;  // query rows of columns (2t, 2t+1)
          const float l0 = s_lse[rr], l1 = s_lse[rr + 1], d0 = s_delta[rr], d1 = s_delta[rr + 1];
          const float p00 = exp2f(st[a][nt][0] * sl2 - l0), p01 = exp2f(st[a][nt][1] * sl2 - l1);
          const float p10 = exp2f(st[a][nt][2] * sl2 - l0), p11 = exp2f(st[a][nt][3] * sl2 - l1);
          const float s00 = p00 * (dpt[a][nt][0] - d0), s01 = p01 * (dpt[a][nt][1] - d1);
          const float s10 = p10 * (dpt[a][nt][2] - d0), s11 = p11 * (dpt[a][nt][3] - d1);
          pf[a][nt >> 1][(nt & 1) * 2 + 0] = ab_pack(p00, p01);
          pf[a][nt >> 1][(nt & 1) * 2 + 1] = ab_pack(p10, p11);
          dsf[a][nt >> 1][(nt & 1) * 2 + 0] = ab_pack(s00, s01);
          dsf[a][nt >> 1][(nt & 1) * 2 + 1] = ab_pack(s10, s11);
          // dS^T to smem for the dQ product (keys x 32 rows): thread holds (key g / g+8, rows 2t,2t+1) of n-tile nt
          const int k_lo = key0 + a * 16 + g;
          *reinterpret_cast<uint32_t*>(ab_ds(sDS, k_lo, nt) + t4 * 2) = ab_pack(s00, s01);
          *reinterpret_cast<uint32_t*>(ab_ds(sDS, k_lo + 8, nt) + t4 * 2) = ab_pack(s10, s11);
        }
      // dV_w += P^T dO_b ; dK_w += dS^T Q_b : M = 32 keys, N = 64 d, K = 32 rows
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t gf[4], qf[4];
          const int r = qb + ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          ab_ldsm_x4_t(gf, ab_tile(sDO, r, dp * 2 + (lane >> 4)));
          ab_ldsm_x4_t(qf, ab_tile(sQ, r, dp * 2 + (lane >> 4)));
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            ab_mma(dV[a][2 * dp], pf[a][ks], gf[0], gf[1]);
            ab_mma(dV[a][2 * dp + 1], pf[a][ks], gf[2], gf[3]);
            ab_mma(dK[a][2 * dp], dsf[a][ks], qf[0], qf[1]);
            ab_mma(dK[a][2 * dp + 1], dsf[a][ks], qf[2], qf[3]);
          }
        }
    }
    __syncthreads();  // dS^T of all key slices visible
    {
      // dQ_b = dS K : warp -> rows (warp & 1) * 16, d columns (warp >> 1) * 16 ; M = 16, N = 16, K = n_tok keys
      const int mrow = (warp & 1) * 16, dcol = (warp >> 1) * 16;
      float dq[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
      for (int kk = 0; kk < n_tok; kk += 16) {
        // A = dS[row][key] from the transposed store [key][row]: ldmatrix.trans, matrices (keys 0-7|8-15) x (rows 0-7|8-15)
        uint32_t af[4], kf[4];
        const int key = kk + (lane & 7) + (lane >> 4) * 8;
        const int rchunk = (mrow >> 3) + ((lane >> 3) & 1);
        ab_ldsm_x4_t(af, ab_ds(sDS, key, rchunk));
        // B = K[key][d] row-major -> ldmatrix.trans (as V in the forward kernel)
        ab_ldsm_x4_t(kf, ab_tile(sK, kk + (lane & 7) + ((lane >> 3) & 1) * 8, (dcol >> 3) + (lane >> 4)));
        ab_mma(dq[0], af, kf[0], kf[1]);
        ab_mma(dq[1], af, kf[2], kf[3]);
      }
      // dq = scale * dS K
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        bf16* o0 = dqkv + (row0 + qb + mrow + g) * ld + head * AB_HD + dcol + i * 8 + t4 * 2;
        *reinterpret_cast<uint32_t*>(o0) = ab_pack(dq[i][0] * 0.125f, dq[i][1] * 0.125f);
        *reinterpret_cast<uint32_t*>(o0 + 8 * ld) = ab_pack(dq[i][2] * 0.125f, dq[i][3] * 0.125f);
      }
    }
    __syncthreads();  // sDS is rewritten by the next query block
  }
  if (kv_active) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long r = row0 + key0 + a * 16 + g;
        bf16* ko = dqkv + r * ld + D + head * AB_HD + i * 8 + t4 * 2;
        bf16* vo = ko + D;
        *reinterpret_cast<uint32_t*>(ko) = ab_pack(dK[a][i][0] * 0.125f, dK[a][i][1] * 0.125f);
        *reinterpret_cast<uint32_t*>(ko + 8 * ld) = ab_pack(dK[a][i][2] * 0.125f, dK[a][i][3] * 0.125f);
        *reinterpret_cast<uint32_t*>(vo) = ab_pack(dV[a][i][0], dV[a][i][1]);
        *reinterpret_cast<uint32_t*>(vo + 8 * ld) = ab_pack(dV[a][i][2], dV[a][i][3]);
      }
  }
}


// ================================================================================================================
// More than 256 tokens per sample (the 512- / 1024-px models, tld/train.py:166 trains any image_size): the same mma.sync
// scheme tiled over KEY blocks of 256.
//   attention_bwd_stats_kernel   CTA = (head, sample, 256 query rows): delta[r] = dO[r].O[r] and the row log-sum-exp over ALL
//                                keys (K streamed through shared memory in 256-key chunks, running max / sum in registers)
//   attention_bwd_tiled_kernel   CTA = (head, sample, 256-key block): K / V of the block stay in shared memory, the query rows
//                                stream through in chunks of 256; dK / dV of the block accumulate in registers over all
//                                chunks; the block's contribution to dQ is added to an fp32 buffer with red.global.add (the
//                                sum over key blocks; summation order = arrival order, so dQ is not bitwise reproducible
//                                beyond 256 tokens - the <= 256-token kernel above stays deterministic)
//   attention_bwd_dq_cast_kernel dq fp32 * 1/8 -> bf16 into the q columns of dqkv
// ================================================================================================================
__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_stats_kernel(const bf16* __restrict__ qkv, const float* __restrict__ d_out, const float* __restrict__ x_before,
                           const float* __restrict__ x_after, float* __restrict__ lse_g, float* __restrict__ delta_g,
                           int n_tok, int D) {
  extern __shared__ __align__(128) uint8_t ab_smem[];
  bf16* sQ = reinterpret_cast<bf16*>(ab_smem);
  bf16* sK = sQ + AB_MAXN * AB_HD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int head = blockIdx.x, b = blockIdx.y, q0 = blockIdx.z * AB_MAXN;
  const int H = D / AB_HD;
  const long long ld = 3LL * D;
  const long long row0 = (long long)b * n_tok;
  const float sl2 = 0.125f * 1.4426950408889634f;
  float* lse_o = lse_g + ((size_t)b * H + head) * n_tok + q0;
  float* delta_o = delta_g + ((size_t)b * H + head) * n_tok + q0;
  for (int i = tid; i < AB_MAXN * 8; i += AB_THREADS) {
    const int r = i >> 3, c = i & 7;
    const bf16* src = qkv + (row0 + q0 + r) * ld + head * AB_HD + c * 8;
    *reinterpret_cast<uint4*>(ab_tile(sQ, r, c)) = *reinterpret_cast<const uint4*>(src);
    const size_t off = (size_t)(row0 + q0 + r) * D + head * AB_HD + c * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(d_out + off), g1 = *reinterpret_cast<const float4*>(d_out + off + 4);
    const float4 a0 = *reinterpret_cast<const float4*>(x_after + off), a1 = *reinterpret_cast<const float4*>(x_after + off + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(x_before + off), b1 = *reinterpret_cast<const float4*>(x_before + off + 4);
    float part = g0.x * (a0.x - b0.x) + g0.y * (a0.y - b0.y) + g0.z * (a0.z - b0.z) + g0.w * (a0.w - b0.w) +
                 g1.x * (a1.x - b1.x) + g1.y * (a1.y - b1.y) + g1.z * (a1.z - b1.z) + g1.w * (a1.w - b1.w);
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    if (c == 0) delta_o[r] = part;
  }
  __syncthreads();
  uint32_t qf[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ab_ldsm_x4(qf[mt][ks], ab_tile(sQ, warp * 32 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4)));
  float m_run[2][2] = {{-INFINITY, -INFINITY}, {-INFINITY, -INFINITY}}, l_run[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int kc0 = 0; kc0 < n_tok; kc0 += AB_MAXN) {
    __syncthreads();   // the previous chunk of K has been consumed
    for (int i = tid; i < AB_MAXN * 8; i += AB_THREADS) {
      const int r = i >> 3, c = i & 7;
      *reinterpret_cast<uint4*>(ab_tile(sK, r, c)) =
          *reinterpret_cast<const uint4*>(qkv + (row0 + kc0 + r) * ld + D + head * AB_HD + c * 8);
    }
    __syncthreads();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      for (int kc = 0; kc < AB_MAXN; kc += 64) {
        float sc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            uint32_t kf[4];
            ab_ldsm_x4(kf, ab_tile(sK, kc + np * 16 + (lane & 7) + (lane >> 4) * 8, ks * 2 + ((lane >> 3) & 1)));
            ab_mma(sc[2 * np], qf[mt][ks], kf[0], kf[1]);
            ab_mma(sc[2 * np + 1], qf[mt][ks], kf[2], kf[3]);
          }
        float mx[2] = {m_run[mt][0], m_run[mt][1]};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          mx[0] = fmaxf(mx[0], fmaxf(sc[i][0], sc[i][1]));
          mx[1] = fmaxf(mx[1], fmaxf(sc[i][2], sc[i][3]));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
          mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
        float rs[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          rs[0] += exp2f((sc[i][0] - mx[0]) * sl2) + exp2f((sc[i][1] - mx[0]) * sl2);
          rs[1] += exp2f((sc[i][2] - mx[1]) * sl2) + exp2f((sc[i][3] - mx[1]) * sl2);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          l_run[mt][r] = l_run[mt][r] * exp2f((m_run[mt][r] - mx[r]) * sl2) + rs[r];
          m_run[mt][r] = mx[r];
        }
      }
    }
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l_run[mt][r] += __shfl_xor_sync(0xffffffffu, l_run[mt][r], 1);
      l_run[mt][r] += __shfl_xor_sync(0xffffffffu, l_run[mt][r], 2);
    }
    if (t4 == 0) {
      lse_o[warp * 32 + mt * 16 + g] = m_run[mt][0] * sl2 + log2f(l_run[mt][0]);
      lse_o[warp * 32 + mt * 16 + g + 8] = m_run[mt][1] * sl2 + log2f(l_run[mt][1]);
    }
  }
}

__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_tiled_kernel(const bf16* __restrict__ qkv, const float* __restrict__ d_out, const float* __restrict__ lse_g,
                           const float* __restrict__ delta_g, float* __restrict__ dq_acc, bf16* __restrict__ dqkv, int n_tok,
                           int D) {
  extern __shared__ __align__(128) uint8_t ab_smem[];
  bf16* sQ = reinterpret_cast<bf16*>(ab_smem);
  bf16* sK = sQ + AB_MAXN * AB_HD;
  bf16* sV = sK + AB_MAXN * AB_HD;
  bf16* sDO = sV + AB_MAXN * AB_HD;
  bf16* sDS = sDO + AB_MAXN * AB_HD;                       // [256 keys][32 rows]
  float* s_lse = reinterpret_cast<float*>(sDS + AB_MAXN * 32);
  float* s_delta = s_lse + AB_MAXN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int head = blockIdx.x, b = blockIdx.y, k0 = blockIdx.z * AB_MAXN;
  const int H = D / AB_HD;
  const long long ld = 3LL * D;
  const long long row0 = (long long)b * n_tok;
  const float sl2 = 0.125f * 1.4426950408889634f;
  for (int i = tid; i < AB_MAXN * 8; i += AB_THREADS) {
    const int r = i >> 3, c = i & 7;
    const bf16* src = qkv + (row0 + k0 + r) * ld + head * AB_HD + c * 8;
    *reinterpret_cast<uint4*>(ab_tile(sK, r, c)) = *reinterpret_cast<const uint4*>(src + D);
    *reinterpret_cast<uint4*>(ab_tile(sV, r, c)) = *reinterpret_cast<const uint4*>(src + 2 * D);
  }
  const int key0 = warp * 32;              // this warp's key slice inside the block
  float dV[2][8][4], dK[2][8][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dV[a][i][j] = dK[a][i][j] = 0.f;

  for (int qc0 = 0; qc0 < n_tok; qc0 += AB_MAXN) {
    __syncthreads();   // previous chunk fully consumed (also orders the K / V loads before the first use)
    for (int i = tid; i < AB_MAXN * 8; i += AB_THREADS) {
      const int r = i >> 3, c = i & 7;
      *reinterpret_cast<uint4*>(ab_tile(sQ, r, c)) =
          *reinterpret_cast<const uint4*>(qkv + (row0 + qc0 + r) * ld + head * AB_HD + c * 8);
      const float* go = d_out + (row0 + qc0 + r) * D + head * AB_HD + c * 8;
      const float4 g0 = *reinterpret_cast<const float4*>(go), g1 = *reinterpret_cast<const float4*>(go + 4);
      uint4 pk;
      pk.x = ab_pack(g0.x, g0.y); pk.y = ab_pack(g0.z, g0.w); pk.z = ab_pack(g1.x, g1.y); pk.w = ab_pack(g1.z, g1.w);
      *reinterpret_cast<uint4*>(ab_tile(sDO, r, c)) = pk;
    }
    if (tid < AB_MAXN) {
      s_lse[tid] = lse_g[((size_t)b * H + head) * n_tok + qc0 + tid];
      s_delta[tid] = delta_g[((size_t)b * H + head) * n_tok + qc0 + tid];
    }
    __syncthreads();
    for (int qb = 0; qb < AB_MAXN; qb += 32) {
      {
        float st[2][4][4], dpt[2][4][4];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) st[a][i][j] = dpt[a][i][j] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t kf[2][4], vf[2][4];
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const int r = key0 + a * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            ab_ldsm_x4(kf[a], ab_tile(sK, r, ks * 2 + (lane >> 4)));
            ab_ldsm_x4(vf[a], ab_tile(sV, r, ks * 2 + (lane >> 4)));
          }
#pragma unroll
          for (int np = 0; np < 2; ++np) {
            uint32_t qf[4], gf[4];
            const int r = qb + np * 16 + (lane & 7) + (lane >> 4) * 8;
            ab_ldsm_x4(qf, ab_tile(sQ, r, ks * 2 + ((lane >> 3) & 1)));
            ab_ldsm_x4(gf, ab_tile(sDO, r, ks * 2 + ((lane >> 3) & 1)));
#pragma unroll
            for (int a = 0; a < 2; ++a) {
              ab_mma(st[a][2 * np], kf[a], qf[0], qf[1]);
              ab_mma(st[a][2 * np + 1], kf[a], qf[2], qf[3]);
              ab_mma(dpt[a][2 * np], vf[a], gf[0], gf[1]);
              ab_mma(dpt[a][2 * np + 1], vf[a], gf[2], gf[3]);
            }
          }
        }
        uint32_t pf[2][2][4], dsf[2][2][4];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const int rr = qb + nt * 8 + t4 * 2;
            const float l0 = s_lse[rr], l1 = s_lse[rr + 1], d0 = s_delta[rr], d1 = s_delta[rr + 1];
            const float p00 = exp2f(st[a][nt][0] * sl2 - l0), p01 = exp2f(st[a][nt][1] * sl2 - l1);
            const float p10 = exp2f(st[a][nt][2] * sl2 - l0), p11 = exp2f(st[a][nt][3] * sl2 - l1);
            const float s00 = p00 * (dpt[a][nt][0] - d0), s01 = p01 * (dpt[a][nt][1] - d1);
            const float s10 = p10 * (dpt[a][nt][2] - d0), s11 = p11 * (dpt[a][nt][3] - d1);
            pf[a][nt >> 1][(nt & 1) * 2 + 0] = ab_pack(p00, p01);
            pf[a][nt >> 1][(nt & 1) * 2 + 1] = ab_pack(p10, p11);
            dsf[a][nt >> 1][(nt & 1) * 2 + 0] = ab_pack(s00, s01);
            dsf[a][nt >> 1][(nt & 1) * 2 + 1] = ab_pack(s10, s11);
            const int k_lo = key0 + a * 16 + g;
            *reinterpret_cast<uint32_t*>(ab_ds(sDS, k_lo, nt) + t4 * 2) = ab_pack(s00, s01);
            *reinterpret_cast<uint32_t*>(ab_ds(sDS, k_lo + 8, nt) + t4 * 2) = ab_pack(s10, s11);
          }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            uint32_t gf[4], qf[4];
            const int r = qb + ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            ab_ldsm_x4_t(gf, ab_tile(sDO, r, dp * 2 + (lane >> 4)));
            ab_ldsm_x4_t(qf, ab_tile(sQ, r, dp * 2 + (lane >> 4)));
#pragma unroll
            for (int a = 0; a < 2; ++a) {
              ab_mma(dV[a][2 * dp], pf[a][ks], gf[0], gf[1]);
              ab_mma(dV[a][2 * dp + 1], pf[a][ks], gf[2], gf[3]);
              ab_mma(dK[a][2 * dp], dsf[a][ks], qf[0], qf[1]);
              ab_mma(dK[a][2 * dp + 1], dsf[a][ks], qf[2], qf[3]);
            }
          }
      }
      __syncthreads();  // dS^T of all key slices visible
      {
        // this key block's share of dQ_b = dS K: warp -> rows (warp & 1) * 16, d columns (warp >> 1) * 16, K = 256 keys
        const int mrow = (warp & 1) * 16, dcol = (warp >> 1) * 16;
        float dq[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
        for (int kk = 0; kk < AB_MAXN; kk += 16) {
          uint32_t af[4], kf[4];
          const int key = kk + (lane & 7) + (lane >> 4) * 8;
          const int rchunk = (mrow >> 3) + ((lane >> 3) & 1);
          ab_ldsm_x4_t(af, ab_ds(sDS, key, rchunk));
          ab_ldsm_x4_t(kf, ab_tile(sK, kk + (lane & 7) + ((lane >> 3) & 1) * 8, (dcol >> 3) + (lane >> 4)));
          ab_mma(dq[0], af, kf[0], kf[1]);
          ab_mma(dq[1], af, kf[2], kf[3]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float* o0 = dq_acc + (size_t)(row0 + qc0 + qb + mrow + g) * D + head * AB_HD + dcol + i * 8 + t4 * 2;
          atomicAdd(o0, dq[i][0]);
          atomicAdd(o0 + 1, dq[i][1]);
          atomicAdd(o0 + (size_t)8 * D, dq[i][2]);
          atomicAdd(o0 + (size_t)8 * D + 1, dq[i][3]);
        }
      }
      __syncthreads();  // sDS is rewritten by the next query block
    }
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long r = row0 + k0 + key0 + a * 16 + g;
      bf16* ko = dqkv + r * ld + D + head * AB_HD + i * 8 + t4 * 2;
      bf16* vo = ko + D;
      *reinterpret_cast<uint32_t*>(ko) = ab_pack(dK[a][i][0] * 0.125f, dK[a][i][1] * 0.125f);
      *reinterpret_cast<uint32_t*>(ko + 8 * ld) = ab_pack(dK[a][i][2] * 0.125f, dK[a][i][3] * 0.125f);
      *reinterpret_cast<uint32_t*>(vo) = ab_pack(dV[a][i][0], dV[a][i][1]);
      *reinterpret_cast<uint32_t*>(vo + 8 * ld) = ab_pack(dV[a][i][2], dV[a][i][3]);
    }
}

__global__ void __launch_bounds__(256) attention_bwd_dq_cast_kernel(const float* __restrict__ dq_acc, bf16* __restrict__ dqkv,
                                                                    long long T, int D) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;   // one thread = 4 columns
  const int d4 = D / 4;
  if (i >= T * d4) return;
  const long long r = i / d4;
  const int c = int(i % d4) * 4;
  const float4 v = *reinterpret_cast<const float4*>(dq_acc + r * D + c);
  uint2 o;
  o.x = ab_pack(v.x * 0.125f, v.y * 0.125f);
  o.y = ab_pack(v.z * 0.125f, v.w * 0.125f);
  *reinterpret_cast<uint2*>(dqkv + r * 3LL * D + c) = o;
}

// 0 / 1: the mma.sync kernels (default); 2: the tcgen05 kernel of attention_bwd_tc.cu when n_tok % 256 == 0.  Measured on B200
// (tools/sab_probe.py, one call = all kernels of the backward of one layer's attention): 256 tokens x batch 32: 128 us (mma.sync)
// vs 205 us (tcgen05 133 + statistics 52 + cast 8 + memset); 1024 tokens x 8: 441 vs 602 us; 4096 x 2: 1517 vs 2186 us.  The
// tensor-core version is a serial chain per CTA (TMA -> MMA -> tcgen05.ld -> softmax -> st.shared -> MMA -> atomics) with one CTA
// per SM (448 TMEM columns), whereas the mma.sync kernel keeps every intermediate in registers and overlaps 8 warps; it needs two
// query tiles in flight per CTA to win, which 512 TMEM columns do not allow with this tiling.  Kept selectable and parity-tested.
static int g_attn_bwd_impl = 0;
void set_attention_bwd_impl(int v) { g_attn_bwd_impl = v; }
int launch_self_attention_bwd_tc(const bf16* qkv, const float* d_out, const float* lse_g, const float* delta_g, float* dq_acc,
                                 bf16* dqkv, int B, int n_tok, int D, cudaStream_t st);

int launch_self_attention_bwd(const bf16* qkv, const float* d_out, const float* x_before, const float* x_after, bf16* dqkv,
                              int B, int n_tok, int D, cudaStream_t st) {
  TLD_CHECK(D % 64 == 0 && n_tok % 64 == 0 && (n_tok <= AB_MAXN || n_tok % AB_MAXN == 0),
            "self_attention_bwd: needs embed_dim % 64 == 0 and tokens per sample in {64,128,192,256} or a multiple of 256");
  TLD_CHECK(B <= 65535, "self_attention_bwd: batch too large");
  constexpr int smem = (4 * AB_MAXN * AB_HD + AB_MAXN * 32) * 2 + 2 * AB_MAXN * 4;
  const bool use_tc = g_attn_bwd_impl == 2 && n_tok % AB_MAXN == 0;
  if (n_tok > AB_MAXN || use_tc) {
    // key-tiled paths: statistics pass, then either the tcgen05 kernel (one CTA per 128-key block) or the mma.sync kernel (one CTA
    // per 256-key block), both accumulating dQ into an fp32 buffer; then the cast
    const long long T = (long long)B * n_tok;
    const int H = D / 64, nblk = n_tok / AB_MAXN;
    TLD_CHECK(nblk <= 65535, "self_attention_bwd: too many tokens per sample");
    float* scr = device_scratch(SCR_ATTN_BWD, (size_t)T * D + 2 * (size_t)T * H);
    if (!scr) return 1;
    float *dq_acc = scr, *lse_g = scr + (size_t)T * D, *delta_g = lse_g + (size_t)T * H;
    static bool attr2 = false;
    if (!attr2) {
      TLD_CUDA_OK(cudaFuncSetAttribute(attention_bwd_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * AB_MAXN * AB_HD * 2));
      TLD_CUDA_OK(cudaFuncSetAttribute(attention_bwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr2 = true;
    }
    TLD_CUDA_OK(cudaMemsetAsync(dq_acc, 0, sizeof(float) * (size_t)T * D, st));
    attention_bwd_stats_kernel<<<dim3(H, B, nblk), AB_THREADS, 2 * AB_MAXN * AB_HD * 2, st>>>(qkv, d_out, x_before, x_after, lse_g,
                                                                                            delta_g, n_tok, D);
    TLD_CUDA_OK(cudaGetLastError());
    if (use_tc) {
      if (launch_self_attention_bwd_tc(qkv, d_out, lse_g, delta_g, dq_acc, dqkv, B, n_tok, D, st)) return 1;
    } else {
      attention_bwd_tiled_kernel<<<dim3(H, B, nblk), AB_THREADS, smem, st>>>(qkv, d_out, lse_g, delta_g, dq_acc, dqkv, n_tok, D);
      TLD_CUDA_OK(cudaGetLastError());
    }
    attention_bwd_dq_cast_kernel<<<(unsigned)((T * (D / 4) + 255) / 256), 256, 0, st>>>(dq_acc, dqkv, T, D);
    TLD_CUDA_OK(cudaGetLastError());
    return 0;
  }
  static bool attr_set = false;
  if (!attr_set) {
    TLD_CUDA_OK(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  attention_bwd_kernel<<<dim3(D / 64, B), AB_THREADS, smem, st>>>(qkv, d_out, x_before, x_after, dqkv, n_tok, D);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace tld

extern "C" __attribute__((visibility("default"))) int tld_bwd_self_attention(const uint16_t* qkv, const float* d_out,
                                                                             const float* x_before, const float* x_after,
                                                                             uint16_t* dqkv, int batch, int n_tok, int D,
                                                                             void* stream) {
  return tld::launch_self_attention_bwd(reinterpret_cast<const tld::bf16*>(qkv), d_out, x_before, x_after,
                                        reinterpret_cast<tld::bf16*>(dqkv), batch, n_tok, D,
                                        reinterpret_cast<cudaStream_t>(stream));
}
