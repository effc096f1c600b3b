// tcgen05 flash-attention for head_dim 64, non-causal, fused with the residual add (no out-proj in the reference):
//     x[t, h*64:(h+1)*64] += softmax(q k^T / 8) v            (transformer_blocks.py:31-48,57-59,136)
//
// CTA = 128 query rows of one (sample, head); keys are consumed in chunks of 128:
//     S[128x128] = Q K_c^T          tcgen05.mma, A = Q (smem, K-major), B = K_c (smem, K-major), D in TMEM cols [0,128)
//     P = exp2(S*c - m*c)           256 softmax threads, two per row (64 keys each): tcgen05.ld -> max/sum in registers,
//                                   P written as bf16 into 128B-swizzled K-major smem tiles
//     O_c[128x64] = P V_c           tcgen05.mma, A = P (smem, K-major), B = V_c (smem, MN-major: hd contiguous),
//                                   D in TMEM cols [128,192); added into the register accumulator with the online-
//                                   softmax correction, so TMEM never needs rescaling
// Output: O/l -> swizzled fp32 staging -> TMA reduce-add into the fp32 residual stream.
// Warps 0..7: softmax + epilogue, two threads per query row (TMEM lane quarter = warp & 3).  Warp 8: TMEM alloc, TMA
// loads, MMA issue.
// 80 KB smem + 256 TMEM columns per CTA -> two CTAs per SM overlap each other's MMA and MUFU phases.
#include "common.h"
#include "launch.h"
#include "ptx.cuh"

namespace tld {

constexpr int TA_BQ = 128, TA_BK = 128, TA_HD = 64, TA_THREADS = 288;  // 8 softmax warps + 1 control warp
constexpr int TA_Q_BYTES = TA_BQ * TA_HD * 2;   // 16 KB
constexpr int TA_K_BYTES = TA_BK * TA_HD * 2;   // 16 KB
constexpr int TA_P_BYTES = TA_BQ * TA_BK * 2;   // 32 KB (two [128 x 64] K-major sub-tiles); reused as fp32 staging
constexpr int TA_SMEM = 1024 + TA_Q_BYTES + 2 * TA_K_BYTES + TA_P_BYTES + 128 + 2 * 128 * 4;  // + max/sum exchange

__global__ void __launch_bounds__(TA_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_x, int n_tok,
                    int D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TA_Q_BYTES;
  uint8_t* sV = sK + TA_K_BYTES;
  uint8_t* sP = sV + TA_K_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + TA_P_BYTES);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_p = bars + 4;
  uint64_t* bar_o = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* s_xchg = reinterpret_cast<float*>(bars + 16);  // [2][128] row max / row sum exchange between the two halves

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int n_chunks = n_tok / TA_BK;
  const int row_q = b * n_tok + qt * TA_BQ;   // first query row (token index) of this CTA
  const int row_k = b * n_tok;                // first key row of this sample
  const int col_q = head * TA_HD, col_k = D + head * TA_HD, col_v = 2 * D + head * TA_HD;

  pdl_launch_dependents();
  if (warp == 8) {
    if (elect_one()) {
      tma_prefetch_desc(&tmap_qkv);
      tma_prefetch_desc(&tmap_x);
      mbar_init(bar_q, 1);
      mbar_init(bar_k, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 256);
      mbar_init(bar_o, 1);
      fence_mbar_init();
      pdl_wait();  // qkv is the previous kernel's output
      // first loads go out before the TMEM allocation / CTA-wide sync
      mbar_expect_tx(bar_q, TA_Q_BYTES);
      tma_load_2d(sQ, &tmap_qkv, bar_q, col_q, row_q);
      mbar_expect_tx(bar_k, TA_K_BYTES);
      tma_load_2d(sK, &tmap_qkv, bar_k, col_k, row_k);
      mbar_expect_tx(bar_v, TA_K_BYTES);
      tma_load_2d(sV, &tmap_qkv, bar_v, col_v, row_k);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // every thread (the epilogue updates x, which earlier kernels read)
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;         // S: columns [0,128)
  const uint32_t tmem_o = tmem_base + 128;   // O chunk: columns [128,192)

  if (warp == 8) {
    // ===================== control: TMA loads + MMA issue (one lane) =====================
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(TA_BQ, TA_BK, 0, 0);   // S = Q K^T : both K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(TA_BQ, TA_HD, 0, 1);   // O = P V   : B (V) is MN-major
      const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(sK), 16, 1024);
      auto issue_s = [&]() {
#pragma unroll
        for (int k = 0; k < TA_HD / 16; ++k) umma_ss_f16(tmem_s, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        umma_commit(bar_s);
      };
      mbar_wait(bar_q, 0);
      mbar_wait(bar_k, 0);
      tc_fence_after();
      issue_s();
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t ph = c & 1;
        const bool more = c + 1 < n_chunks;
        mbar_wait(bar_s, ph);  // S(c) complete -> the K buffer is free
        if (more) {
          mbar_expect_tx(bar_k, TA_K_BYTES);
          tma_load_2d(sK, &tmap_qkv, bar_k, col_k, row_k + (c + 1) * TA_BK);
        }
        mbar_wait(bar_p, ph);  // P(c) written and S(c) fully read by all 256 softmax threads
        mbar_wait(bar_v, ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < TA_BK / 16; ++kk) {
          // A: P sub-tile kk/4 ([128 x 64] K-major), 32 B per k-step inside the swizzle row
          const uint64_t adesc = umma_smem_desc_sw128(smem_u32(sP + (kk >> 2) * (TA_BQ * 128)), 16, 1024) + 2 * (kk & 3);
          // B: V rows kk*16.. (keys) x 64 hd, MN-major: 8-key groups are 1024 B apart (SBO)
          const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(sV + kk * 16 * 128), TA_BK * 128, 1024);
          umma_ss_f16(tmem_o, adesc, bdesc, idesc_o, kk != 0);
        }
        umma_commit(bar_o);
        if (more) {  // S(c+1) runs behind PV(c) on the tensor pipe while the softmax threads read O(c)
          mbar_wait(bar_k, ph ^ 1);
          tc_fence_after();
          issue_s();
        }
        mbar_wait(bar_o, ph);  // O(c) complete -> V and P buffers are free
        if (more) {
          mbar_expect_tx(bar_v, TA_K_BYTES);
          tma_load_2d(sV, &tmap_qkv, bar_v, col_v, row_k + (c + 1) * TA_BK);
        }
      }
    }
  } else {
    // ===================== softmax + epilogue: TWO threads per query row =====================
    // thread (hf, r): row r = tid & 127, half hf = tid >> 7 owns keys [64 hf, 64 hf + 64) of every chunk (= P sub-tile
    // hf) and output columns [32 hf, 32 hf + 32).  Warps w and w+4 share TMEM lane quarter w & 3.  The row maximum is
    // exchanged through shared memory once per chunk (one 256-thread named barrier), the row sums once at the end.
    const int r = threadIdx.x & 127, hf = threadIdx.x >> 7;
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    const float sl2 = 0.125f * 1.4426950408889634f;
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
      const uint32_t ph = c & 1;
      mbar_wait(bar_s, ph);
      tc_fence_after();
      // pass 1: maximum over this thread's 64 keys, then combine with the partner thread
      float mloc = -INFINITY;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t s[32];
        tmem_ld_x32(tmem_s + lane_base + hf * 64 + j * 32, s);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) mloc = fmaxf(mloc, __uint_as_float(s[i]));
      }
      s_xchg[hf * 128 + r] = mloc;
      named_bar_sync(1, 256);
      const float mx = fmaxf(m_run, fmaxf(mloc, s_xchg[(hf ^ 1) * 128 + r]));
      const float corr = exp2f((m_run - mx) * sl2);
      const float mb = mx * sl2;
      m_run = mx;
      // pass 2: P = exp2(s*c - m*c) -> bf16 -> swizzled K-major smem (sub-tile hf); partial row sum
      float2 rs2 = make_float2(0.f, 0.f);
      uint8_t* sub = sP + hf * (TA_BQ * 128) + r * 128;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t s[32];
        tmem_ld_x32(tmem_s + lane_base + hf * 64 + j * 32, s);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 a = ffma2(make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])),
                                 make_float2(sl2, sl2), make_float2(-mb, -mb));
          const float2 p = make_float2(ex2_approx(a.x), ex2_approx(a.y));
          rs2 = fadd2(rs2, p);
          pk[i] = pack_bf16x2(p.x, p.y);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = j * 4 + q;
          *reinterpret_cast<uint4*>(sub + ((chunk ^ (r & 7)) << 4)) =
              make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      l_run = l_run * corr + (rs2.x + rs2.y);
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(bar_p);
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] *= corr;
      mbar_wait(bar_o, ph);
      tc_fence_after();
      {
        uint32_t v[32];
        tmem_ld_x32(tmem_o + lane_base + hf * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] += __uint_as_float(v[i]);
      }
      tc_fence_before();
    }
    // epilogue: x += O / l.  Row sum = own half + partner's half; each warp stages its [32 rows x 32 cols] fp32 block in
    // its own 4 KB slab (aliasing the dead P buffer) and issues one TMA reduce-add.
    named_bar_sync(1, 256);  // everyone is past the last use of s_xchg (and of P: bar_o of the last chunk)
    s_xchg[hf * 128 + r] = l_run;
    named_bar_sync(1, 256);
    const float inv = 1.f / (l_run + s_xchg[(hf ^ 1) * 128 + r]);
    uint8_t* slab = sP + warp * (32 * 128);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 q4 = make_float4(o[4 * j] * inv, o[4 * j + 1] * inv, o[4 * j + 2] * inv, o[4 * j + 3] * inv);
      *reinterpret_cast<float4*>(slab + lane * 128 + ((j ^ (lane & 7)) << 4)) = q4;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_reduce_add_2d(&tmap_x, slab, col_q + hf * 32, row_q + (warp & 3) * 32);
      bulk_commit();
      bulk_wait_read<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int launch_self_attention_tc(const bf16* qkv, float* x, int B, int n_tok, int D, cudaStream_t st) {
  TLD_CHECK(D % 64 == 0 && n_tok % 128 == 0, "attention_tc: needs embed_dim % 64 == 0 and tokens per sample % 128 == 0");
  TLD_CHECK(B <= 65535, "attention_tc: batch too large for gridDim.z");
  static bool attr_set = false;
  if (!attr_set) {
    TLD_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
    attr_set = true;
  }
  const long long T = (long long)B * n_tok;
  CUtensorMap tq, tx;
  if (make_tmap_2d(&tq, qkv, false, T, 3LL * D, 3LL * D, 128)) return 1;
  if (make_tmap_2d(&tx, x, true, T, D, D, 32)) return 1;
  dim3 grid(n_tok / TA_BQ, D / 64, B);
  return launch_pdl(attention_tc_kernel, grid, dim3(TA_THREADS), TA_SMEM, st, tq, tx, n_tok, D);
}

}  // namespace tld
