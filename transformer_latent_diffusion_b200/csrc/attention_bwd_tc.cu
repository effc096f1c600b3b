// Self-attention backward on the 5th-generation tensor cores (head_dim 64, non-causal); reference: autograd through
// F.scaled_dot_product_attention at tld/transformer_blocks.py:37-44, driven by loss.backward() at tld/train.py:169.
//
// One CTA = one (sample, head, block of 128 keys).  K_j and V_j stay in shared memory; the query rows stream through in tiles of
// 128.  Per (query tile i, key block j) the tensor core runs five products, all operands in 128B-swizzled shared memory, all
// accumulators in TMEM (448 of 512 columns):
//     S  = Q_i K_j^T          M=128 q    N=128 keys K=64    A = Q_i  K-major,        B = K_j K-major          cols [  0,128)
//     dP = dO_i V_j^T         M=128 q    N=128 keys K=64    A = dO_i K-major,        B = V_j K-major          cols [128,256)
//   the 128 softmax threads (thread = query row): P = exp2(S c - lse), dS = P (dP - delta) -> bf16 -> shared memory, stored
//   ONCE as two K-major [128 q x 64 keys] tiles; the same bytes serve as K-major A (dS K) and as MN-major A (P^T, dS^T):
//     dV_j += P^T  dO_i       M=128 keys N=64 d     K=128 q   A = P   MN-major,        B = dO_i MN-major        cols [256,320)
//     dK_j += dS^T Q_i        M=128 keys N=64 d     K=128 q   A = dS  MN-major,        B = Q_i  MN-major        cols [320,384)
//     dQ_i  = dS   K_j        M=128 q    N=64 d     K=128 k   A = dS  K-major,         B = K_j  MN-major        cols [384,448)
//   (MN-major B = the natural [rows][64] tile read transposed, as V in the forward kernel - no transposed copy of anything).
// dV_j / dK_j accumulate in TMEM over all query tiles; dQ_i is this key block's share and is added to an fp32 buffer
// (red.global.add: with the two key blocks of the 256-token model the sum is still order-independent, i.e. deterministic).
// Row statistics (log-sum-exp over ALL keys, delta = dO . O) come from attention_bwd_stats_kernel (attention_bwd.cu).
// The mma.sync kernel this replaces took 128 us per layer at batch 32 (12.8 % of the training step).
#include "common.h"
#include "ptx.cuh"

namespace tld {

constexpr int BT_THREADS = 192;                 // warps 0..3: softmax / epilogue (thread = row), warp 4: TMA + MMA issue
constexpr int BT_TILE = 128 * 64 * 2;           // one [128 x 64] bf16 operand tile: 16 KB
constexpr int BT_SMEM = 1024 + 4 * BT_TILE + 2 * 2 * BT_TILE + 256;   // Q, K, V, dO + P (2 key blocks of 64) + dS

__global__ void __launch_bounds__(BT_THREADS, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const float* __restrict__ d_out,
                        const float* __restrict__ lse_g, const float* __restrict__ delta_g, float* __restrict__ dq_acc,
                        bf16* __restrict__ dqkv, int n_tok, int D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + BT_TILE;
  uint8_t* sV = sK + BT_TILE;
  uint8_t* sDO = sV + BT_TILE;
  uint8_t* sP = sDO + BT_TILE;       // [2 key blocks of 64][128 q][128 B]
  uint8_t* sDS = sP + 2 * BT_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + 2 * BT_TILE);
  uint64_t* bar_kv = bars + 0;
  uint64_t* bar_q = bars + 1;
  uint64_t* bar_do = bars + 2;     // 128 arrivals: dO_i converted into shared memory
  uint64_t* bar_sdp = bars + 3;    // S and dP complete
  uint64_t* bar_pds = bars + 4;    // 128 arrivals: P and dS written
  uint64_t* bar_mma2 = bars + 5;   // dV / dK / dQ products complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int H = D / 64, q_tiles = n_tok / 128;
  const long long row0 = (long long)b * n_tok;
  const float sl2 = 0.125f * 1.4426950408889634f;

  if (warp == 4) {
    if (elect_one()) {
      tma_prefetch_desc(&tmap_qkv);
      mbar_init(bar_kv, 1);
      mbar_init(bar_q, 1);
      mbar_init(bar_do, 128);
      mbar_init(bar_sdp, 1);
      mbar_init(bar_pds, 128);
      mbar_init(bar_mma2, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_s = tmem_base, tm_dp = tmem_base + 128, tm_dv = tmem_base + 256, tm_dk = tmem_base + 320, tm_dq = tmem_base + 384;

  if (warp == 4) {
    // ===================== TMA + MMA issue (one thread) =====================
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);    // S, dP: both operands K-major
      constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64, 1, 1);     // dV, dK: A and B MN-major
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64, 0, 1);     // dQ: A K-major, B MN-major
      mbar_expect_tx(bar_kv, 2 * BT_TILE);
      tma_load_2d(sK, &tmap_qkv, bar_kv, D + head * 64, int(row0) + j * 128);
      tma_load_2d(sV, &tmap_qkv, bar_kv, 2 * D + head * 64, int(row0) + j * 128);
      for (int i = 0; i < q_tiles; ++i) {
        const uint32_t ph = i & 1;
        mbar_expect_tx(bar_q, BT_TILE);   // the previous tile's products (readers of sQ) have retired: bar_mma2 waited below
        tma_load_2d(sQ, &tmap_qkv, bar_q, head * 64, int(row0) + i * 128);
        if (i == 0) mbar_wait(bar_kv, 0);
        mbar_wait(bar_q, ph);
        mbar_wait(bar_do, ph);
        tc_fence_after();
        {
          const uint64_t qd = umma_smem_desc_sw128(smem_u32(sQ), 16, 1024), kd = umma_smem_desc_sw128(smem_u32(sK), 16, 1024);
          const uint64_t gd = umma_smem_desc_sw128(smem_u32(sDO), 16, 1024), vd = umma_smem_desc_sw128(smem_u32(sV), 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss_f16(tm_s, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss_f16(tm_dp, gd + 2 * k, vd + 2 * k, idesc_s, k != 0);
          umma_commit(bar_sdp);
        }
        mbar_wait(bar_pds, ph);
        tc_fence_after();
        {
          // contraction over the 128 query rows: 8 steps of 16 rows = 2048 B; A atoms (64 keys wide) 16 KB apart
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t pa = umma_smem_desc_sw128(smem_u32(sP) + k * 2048, 2 * BT_TILE / 2, 1024);
            const uint64_t gb = umma_smem_desc_sw128(smem_u32(sDO) + k * 2048, 8192, 1024);
            umma_ss_f16(tm_dv, pa, gb, idesc_t, (i | k) != 0);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t sa = umma_smem_desc_sw128(smem_u32(sDS) + k * 2048, 2 * BT_TILE / 2, 1024);
            const uint64_t qb = umma_smem_desc_sw128(smem_u32(sQ) + k * 2048, 8192, 1024);
            umma_ss_f16(tm_dk, sa, qb, idesc_t, (i | k) != 0);
          }
          // contraction over the 128 keys: key block kb = K-major A tile sDS + kb * 16 KB, 4 steps of 32 B inside the swizzle row
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t sa = umma_smem_desc_sw128(smem_u32(sDS) + (k >> 2) * BT_TILE, 16, 1024) + 2 * (k & 3);
            const uint64_t kb = umma_smem_desc_sw128(smem_u32(sK) + k * 2048, 8192, 1024);
            umma_ss_f16(tm_dq, sa, kb, idesc_q, k != 0);
          }
          umma_commit(bar_mma2);
        }
        mbar_wait(bar_mma2, ph);   // sQ may be reloaded, the softmax threads may overwrite sDO / sP / sDS
      }
    }
  } else if (warp < 4) {
    // ===================== softmax backward + epilogues, thread = row =====================
    const int r = threadIdx.x;                       // query row inside the tile / key row inside the block
    const uint32_t lane_base = uint32_t(warp * 32) << 16;
    const uint32_t row_off = r * 128, sw = r & 7;
    for (int i = 0; i < q_tiles; ++i) {
      const uint32_t ph = i & 1;
      const long long qrow = row0 + i * 128 + r;
      // dO_i row: fp32 -> bf16 into the K-major tile (= MN-major B of the dV product)
      const float4* go = reinterpret_cast<const float4*>(d_out + qrow * D + head * 64);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 a = __ldg(go + 2 * c), b4 = __ldg(go + 2 * c + 1);
        sts_v4(smem_u32(sDO) + row_off + ((c ^ sw) << 4), pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b4.x, b4.y),
               pack_bf16x2(b4.z, b4.w));
      }
      const float lse = lse_g[((size_t)b * H + head) * n_tok + i * 128 + r];
      const float delta = delta_g[((size_t)b * H + head) * n_tok + i * 128 + r];
      fence_proxy_async_smem();
      mbar_arrive(bar_do);
      mbar_wait(bar_sdp, ph);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {                  // 32 keys at a time
        uint32_t sv[32], dv[32];
        tmem_ld_x32(tm_s + lane_base + c * 32, sv);
        tmem_ld_x32(tm_dp + lane_base + c * 32, dv);
        tmem_ld_wait();
        uint32_t pp[16], ds[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = exp2f(__uint_as_float(sv[2 * e]) * sl2 - lse), p1 = exp2f(__uint_as_float(sv[2 * e + 1]) * sl2 - lse);
          pp[e] = pack_bf16x2(p0, p1);
          ds[e] = pack_bf16x2(p0 * (__uint_as_float(dv[2 * e]) - delta), p1 * (__uint_as_float(dv[2 * e + 1]) - delta));
        }
        // keys [32 c, 32 c + 32): key block kb = c >> 1, 16-byte chunks 4 (c & 1) .. + 3 of the row
        const uint32_t pb = smem_u32(sP) + (c >> 1) * BT_TILE + row_off, db = smem_u32(sDS) + (c >> 1) * BT_TILE + row_off;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t ch = uint32_t(((c & 1) * 4 + q) ^ sw) << 4;
          sts_v4(pb + ch, pp[4 * q], pp[4 * q + 1], pp[4 * q + 2], pp[4 * q + 3]);
          sts_v4(db + ch, ds[4 * q], ds[4 * q + 1], ds[4 * q + 2], ds[4 * q + 3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(bar_pds);
      mbar_wait(bar_mma2, ph);
      tc_fence_after();
      // this key block's share of dQ_i (the 1/8 is applied by the cast kernel)
      float* dq = dq_acc + qrow * D + head * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_x32(tm_dq + lane_base + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) atomicAdd(dq + c * 32 + e, __uint_as_float(v[e]));
      }
      tc_fence_before();
    }
    // dK_j (x 1/8), dV_j -> bf16 rows of the key block
    const long long krow = row0 + j * 128 + r;
    bf16* ko = dqkv + krow * 3LL * D + D + head * 64;
    bf16* vo = ko + D;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t kk[32], vv[32];
      tmem_ld_x32(tm_dk + lane_base + c * 32, kk);
      tmem_ld_x32(tm_dv + lane_base + c * 32, vv);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 ok, ov;
        ok.x = pack_bf16x2(__uint_as_float(kk[8 * q]) * 0.125f, __uint_as_float(kk[8 * q + 1]) * 0.125f);
        ok.y = pack_bf16x2(__uint_as_float(kk[8 * q + 2]) * 0.125f, __uint_as_float(kk[8 * q + 3]) * 0.125f);
        ok.z = pack_bf16x2(__uint_as_float(kk[8 * q + 4]) * 0.125f, __uint_as_float(kk[8 * q + 5]) * 0.125f);
        ok.w = pack_bf16x2(__uint_as_float(kk[8 * q + 6]) * 0.125f, __uint_as_float(kk[8 * q + 7]) * 0.125f);
        ov.x = pack_bf16x2(__uint_as_float(vv[8 * q]), __uint_as_float(vv[8 * q + 1]));
        ov.y = pack_bf16x2(__uint_as_float(vv[8 * q + 2]), __uint_as_float(vv[8 * q + 3]));
        ov.z = pack_bf16x2(__uint_as_float(vv[8 * q + 4]), __uint_as_float(vv[8 * q + 5]));
        ov.w = pack_bf16x2(__uint_as_float(vv[8 * q + 6]), __uint_as_float(vv[8 * q + 7]));
        *reinterpret_cast<uint4*>(ko + c * 32 + q * 8) = ok;
        *reinterpret_cast<uint4*>(vo + c * 32 + q * 8) = ov;
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_self_attention_bwd_tc(const bf16* qkv, const float* d_out, const float* lse_g, const float* delta_g, float* dq_acc,
                                 bf16* dqkv, int B, int n_tok, int D, cudaStream_t st) {
  TLD_CHECK(D % 64 == 0 && n_tok % 128 == 0 && B <= 65535, "attention_bwd_tc: needs embed_dim % 64 == 0 and tokens % 128 == 0");
  const long long T = (long long)B * n_tok;
  TLD_CHECK(T < (1ll << 31), "attention_bwd_tc: too many rows");
  CUtensorMap tq;
  if (make_tmap_2d(&tq, qkv, false, T, 3LL * D, 3LL * D, 128)) return 1;   // box 128 rows x 64 columns
  static bool set = false;
  if (!set) {
    TLD_CUDA_OK(cudaFuncSetAttribute(attention_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM));
    set = true;
  }
  attention_bwd_tc_kernel<<<dim3(n_tok / 128, D / 64, B), BT_THREADS, BT_SMEM, st>>>(tq, d_out, lse_g, delta_g, dq_acc, dqkv, n_tok, D);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace tld
