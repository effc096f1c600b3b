// MLPSepConv front half as ONE kernel (reference tld/transformer_blocks.py:95-103,110):
//     g = GELU( dwconv3x3( LN3(x) W_up^T + b_up ) + b_dw )                     for the 16 x 16 token grid (256-px model)
//
// The up-projection is the tcgen05 CTA-pair GEMM of gemm_tcgen05.cuh (M = 256 rows per pair tile, BN = 256 channels,
// TMA -> 128B-swizzled smem -> tcgen05.mma.cta_group::2 -> fp32 accumulators in TMEM, two accumulator stages).  With 256
// tokens per sample a pair tile is exactly ONE image, so the 3x3 neighbourhood of every position lives inside the pair:
// CTA 0 owns grid rows 0..7, CTA 1 rows 8..15, and only ONE grid row (16 positions) per 64-channel slab has to cross
// between the two SMs.  Instead of writing the hidden tensor (200 MB per layer at T = 32768) and reading it back in a second
// kernel, the epilogue keeps it on chip:
//
//   warps 4..7   "drain": tcgen05.ld one 64-channel slab of the accumulator (thread = position) -> LayerNorm fold / bias ->
//                bf16 -> this CTA's conv tile in shared memory ([10 grid rows][16][64 ch], 128B-swizzled, rows 1..8 = own
//                rows, row 0 / row 9 = halo); the boundary grid row is also pushed into the PEER CTA's halo row through
//                distributed shared memory (st.async ... mbarrier::complete_tx::bytes: the data and its 2 KB of transaction
//                count land on the peer's cfull barrier, no cluster-scope fence anywhere).
//   warps 8..23  "conv": four groups of 4 warps, group s takes slab s (= conv tile s) of every tile, so the drain runs a
//                whole tile ahead and the conv warps never wait for data; warp j of a group produces grid rows 2j, 2j+1: lane = channel pair, a 4-row x 3-column fp32 window slides along x,
//                packed FFMA2 arithmetic, GELU in its logistic form (2 MUFU ex2 + 2 rcp per channel pair, dwconv_math.cuh:
//                the MUFU unit is idle here and the FMA pipe is the bound), 128-byte coalesced stores of g.
//   warp 0 / 1 / 2   TMA producer / MMA issuer (leader CTA) / TMEM allocator, exactly as in the plain GEMM.
//
// Barriers per CTA: full/empty[stage], tfull/tempty[2] as in the GEMM; cfull[b] (conv tile b written: 4 local drain warps +
// 2 KB of halo bytes from the peer) and cempty[b] (conv tile b consumed: 4 local + 4 peer conv warps - the peer's count because OUR
// boundary drain warp writes into THEIR halo row).  Out-of-image halo rows (row 0 in CTA 0, row 9 in CTA 1) are zeroed once.
//
// LayerNorm fold (optional): with row statistics (sum, sum of squares of the fp32 residual row) from the producer of x and
// gamma folded into the weight, LN3(x) W^T = rstd (x W'^T - mean s) + c with W' = gamma (.) W, s_n = sum_k W'_nk,
// c_n = sum_k beta_k W_nk + b_n: the A operand is bf16(x) itself and the affine map runs on the accumulator.  Without
// statistics the epilogue is acc + c_n (c = bias), bit-identical to EPI_BIAS_BF16.
//
// Roofline: tensor (154.6 GFLOP per layer at T = 32768) with the CUDA-core work riding along: 9 FMA + a 10-op GELU per
// element = 38 packed FFMA2 per 4 outputs per lane -> ~4900 FMA-pipe cycles per 256 x 256 tile against ~6150 tensor cycles.
// Algorithmic bytes per layer: A 50 MB + g 200 MB (the separate kernels moved A 50 + h 200 + h 200 + g 200).
#include <cudaTypedefs.h>

#include "common.h"
#include "dwconv_math.cuh"
#include "launch.h"
#include "ptx.cuh"

namespace tld {

struct FusedUpArgs {
  const float* col_c;      // [N] c_n (bias, or bias + sum_k beta_k W_nk with the LayerNorm fold)
  const float* col_s;      // [N] s_n = sum_k W'_nk (LayerNorm fold) or nullptr
  const float2* row_part;  // [M, n_part] partial (sum, sum of squares) of the fp32 residual row (LayerNorm fold) or nullptr
  int n_part;              // partials per row (K / 32 when a GEMM epilogue produced them)
  float inv_d, ln_eps;     // 1/D and epsilon of the folded LayerNorm
  const float* dw_w9;      // [9, N] depthwise taps, tap-major
  const float* dw_b;       // [N]
  bf16* out;               // [M, N] g
  bf16* hid_out;           // optional [M, N]: the pre-conv hidden tensor as well (training keeps it for the backward)
};

constexpr int FU_BN = 256, FU_BK = 64, FU_STAGES = 4, FU_THREADS = 768;
constexpr int FU_NBUF = 4;                            // conv tiles: one per 64-channel slab of a 256-channel tile
constexpr int FU_A_BYTES = 128 * FU_BK * 2;          // this CTA's 128 rows
constexpr int FU_B_BYTES = (FU_BN / 2) * FU_BK * 2;  // this CTA's half of the W tile
constexpr int FU_STAGE_BYTES = FU_A_BYTES + FU_B_BYTES;
constexpr int FU_TILE_BYTES = 10 * 16 * 128;         // conv tile: 10 grid rows x 16 positions x 64 channels bf16
constexpr int FU_CS_BYTES = 2 * 2 * FU_BN * 4;        // per-column constants c_n, s_n of the tile, double-buffered over tiles
constexpr int FU_SMEM = 1024 + FU_STAGES * FU_STAGE_BYTES + FU_NBUF * FU_TILE_BYTES + FU_CS_BYTES + 256;

// Push 16 bytes into the PEER CTA's shared memory and signal them on the peer's mbarrier (transaction bytes): the
// distributed-shared-memory producer/consumer primitive.  The consumer just waits for its local barrier phase - no
// cluster-scope fence on either side (an acquire.cluster / release.cluster pair compiles to CCTL.IVALL + MEMBAR.GPU per
// use: an L1 invalidate inside the spin loop; measured 2x on the whole kernel).
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, uint32_t cluster_bar, uint32_t a, uint32_t b, uint32_t c,
                                            uint32_t d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(cluster_addr),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(cluster_bar)
               : "memory");
}

__global__ void __launch_bounds__(FU_THREADS, 1)
gemm_up_dwconv_gelu_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M, int N,
                           int K, FusedUpArgs ep) {
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  constexpr int STAGES = FU_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * FU_A_BYTES;
  uint8_t* smem_tile = smem + STAGES * FU_STAGE_BYTES;            // FU_NBUF conv tiles, 1024-aligned
  float* smem_cs = reinterpret_cast<float*>(smem_tile + FU_NBUF * FU_TILE_BYTES);   // [2 tiles][c | s][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_tile + FU_NBUF * FU_TILE_BYTES + FU_CS_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;       // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint64_t* cfull_bar = bars + 2 * STAGES + 4;   // [FU_NBUF]
  uint64_t* cempty_bar = bars + 2 * STAGES + 4 + FU_NBUF;  // [FU_NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4 + 2 * FU_NBUF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = N / FU_BN;
  const int num_tiles = (M / 256) * n_tiles;
  const int k_blocks = (K + FU_BK - 1) / FU_BK;
  const int first_tile = blockIdx.x / 2;
  const int tile_step = gridDim.x / 2;

  pdl_launch_dependents();
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);   // 4 drain warps of each CTA
    }
    for (int s = 0; s < FU_NBUF; ++s) {
      mbar_init(&cfull_bar[s], 4);    // 4 local drain warps (one of them arms the 2 KB of halo the peer pushes: expect_tx)
      mbar_init(&cempty_bar[s], 8);   // the 4 local + 4 peer conv warps that work on slab s
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  // the halo row that lies outside the image: grid row -1 (tile row 0) in CTA 0, grid row 16 (tile row 9) in CTA 1
  if (threadIdx.x < 512) {
    const int b = threadIdx.x >> 7, i = threadIdx.x & 127;   // FU_NBUF buffers x 128 x 16 bytes = 512 threads
    const uint32_t row = leader ? 0u : 9u;
    sts_v4(smem_u32(smem_tile) + b * FU_TILE_BYTES + row * 2048 + i * 16, 0u, 0u, 0u, 0u);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int m0 = (tile / n_tiles) * 256 + int(cta_rank) * 128;
        const int n0 = (tile % n_tiles) * FU_BN + int(cta_rank) * (FU_BN / 2);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full_bar[stage], FU_STAGE_BYTES * 2);
          tma_load_2d_pair(smem_a + stage * FU_A_BYTES, &tmap_a, &full_bar[stage], kb * FU_BK, m0);
          tma_load_2d_pair(smem_b + stage * FU_B_BYTES, &tmap_b, &full_bar[stage], kb * FU_BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_bf16(256, FU_BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * FU_BN;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = umma_smem_desc_sw128(smem_u32(smem_a + stage * FU_A_BYTES), 16u, 1024);
          const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(smem_b + stage * FU_B_BYTES), 16u, 1024);
#pragma unroll
          for (int k = 0; k < FU_BK / 16; ++k) umma_ss_f16_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0);
          umma_commit_pair(&empty_bar[stage]);
          if (kb == k_blocks - 1) umma_commit_pair(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== drain: TMEM -> (LN fold / bias) -> bf16 -> conv tile (+ halo push to the peer) =====================
    const int ew = warp - 4;             // TMEM lane quarter
    const int et = ew * 32 + lane;       // position inside this CTA's 128 (grid row et / 16 of its 8, x = et % 16)
    // the grid row that the peer needs: CTA 0's last row (positions 112..127) -> peer tile row 0; CTA 1's first row -> peer row 9
    const bool push = leader ? (et >= 112) : (et < 16);
    const uint32_t peer = cta_rank ^ 1u;
    const uint32_t peer_row = leader ? 0u : 9u;
    const bool fold = ep.col_s != nullptr;
    int it = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = (tile / n_tiles) * 256 + int(cta_rank) * 128;
      const int n0 = (tile % n_tiles) * FU_BN;
      // per-column constants of this tile -> shared memory (every thread needs every column: broadcast LDS.128 instead of
      // 64 uniform global loads per slab).  Buffer it & 1: the named barrier of tile it + 1 separates its readers from the
      // writers of tile it + 2.
      float* cs = smem_cs + (it & 1) * (2 * FU_BN);
      cs[et] = __ldg(ep.col_c + n0 + et);
      cs[128 + et] = __ldg(ep.col_c + n0 + 128 + et);
      if (fold) {
        cs[FU_BN + et] = __ldg(ep.col_s + n0 + et);
        cs[FU_BN + 128 + et] = __ldg(ep.col_s + n0 + 128 + et);
      }
      float ar = 1.f, br = 0.f;          // value = ar * acc + (br * s_n + c_n)
      if (ep.row_part) {
        const float2* pp = ep.row_part + (size_t)(m0 + et) * ep.n_part;
        float2 pv[32];   // all partials of the row in flight at once (K <= 1024), then a fixed-order (deterministic) sum
#pragma unroll
        for (int u = 0; u < 32; ++u) pv[u] = u < ep.n_part ? __ldg(pp + u) : make_float2(0.f, 0.f);
        float ps = 0.f, pq = 0.f;
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          ps += pv[u].x;
          pq += pv[u].y;
        }
        const float mean = ps * ep.inv_d;
        ar = rsqrtf(fmaxf(pq * ep.inv_d - mean * mean, 0.f) + ep.ln_eps);
        br = -ar * mean;
      }
      named_bar_sync(1, 128);
      const uint32_t cs_u32 = smem_u32(cs);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + acc * FU_BN;
#pragma unroll 1
      for (int s = 0; s < 4; ++s) {                      // slab s -> conv tile s
        mbar_wait(&cempty_bar[s], (uint32_t(it) & 1u) ^ 1u);   // both CTAs' conv warps are done with the previous tile's slab s
        const uint32_t tile_b = smem_u32(smem_tile) + s * FU_TILE_BYTES;
        const uint32_t own = tile_b + (16 + et) * 128;   // tile position 16 + et: swizzle key (16 + et) & 7 == lane & 7
        const int x = et & 15;
        const uint32_t remote = mapa_u32(tile_b + (peer_row * 16 + x) * 128, peer);
        const uint32_t remote_bar = mapa_u32(smem_u32(&cfull_bar[s]), peer);
#pragma unroll
        for (int half = 0; half < 2; ++half) {           // 32 channels at a time keeps the live registers down
          uint32_t ra[32];
          tmem_ld_x32(taddr + s * 64 + half * 32, ra);
          tmem_ld_wait();
          uint32_t o[16];
          const uint32_t cbase = cs_u32 + (s * 64 + half * 32) * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 t = lds_v4(cbase + i * 16);
            if (fold) {
              const float4 sn = lds_v4(cbase + FU_BN * 4 + i * 16);
              t.x = fmaf(br, sn.x, t.x);
              t.y = fmaf(br, sn.y, t.y);
              t.z = fmaf(br, sn.z, t.z);
              t.w = fmaf(br, sn.w, t.w);
            }
            o[2 * i] = pack_bf16x2(fmaf(ar, __uint_as_float(ra[4 * i]), t.x), fmaf(ar, __uint_as_float(ra[4 * i + 1]), t.y));
            o[2 * i + 1] = pack_bf16x2(fmaf(ar, __uint_as_float(ra[4 * i + 2]), t.z), fmaf(ar, __uint_as_float(ra[4 * i + 3]), t.w));
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts_v4(own + (((half * 4 + j) ^ (lane & 7)) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          if (ep.hid_out) {   // training: the backward differentiates through the conv and needs its input
            uint4* hp = reinterpret_cast<uint4*>(ep.hid_out + (size_t)(m0 + et) * N + n0 + s * 64 + half * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) hp[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          }
          if (push) {   // 16 lanes x 128 B per slab = the 2 KB the peer's cfull barrier expects
#pragma unroll
            for (int j = 0; j < 4; ++j)
              st_async_v4(remote + (((half * 4 + j) ^ (x & 7)) << 4), remote_bar, o[4 * j], o[4 * j + 1], o[4 * j + 2],
                          o[4 * j + 3]);
          }
        }
        __syncwarp();
        if (lane == 0) {
          if (ew == 0) mbar_expect_tx(&cfull_bar[s], 16 * 128);   // arrive + this phase's halo bytes from the peer
          else mbar_arrive(&cfull_bar[s]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
    }
  } else if (warp >= 8) {
    // ===================== conv: depthwise 3x3 + bias + GELU from the conv tile, g -> global =====================
    // 16 conv warps: warp (8 + 4 s + j) owns slab s (= conv tile s) of every tile and produces grid rows 2j, 2j+1 of this
    // CTA's 8 (tile rows 2j+1, 2j+2; window rows 2j..2j+3).  4 conv warps per scheduler: the FMA pipe (2 cycles per FFMA2)
    // is the bound of this kernel, and with only 2 per scheduler dependency / dispatch stalls left it half idle.
    const int s = (warp - 8) >> 2;
    const int j = (warp - 8) & 3;
    const uint32_t peer = cta_rank ^ 1u;
    const int lane_chunk = lane >> 2, lane_off = (lane & 3) * 4;
    int it = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
      const int m0 = (tile / n_tiles) * 256 + int(cta_rank) * 128;
      const int n0 = (tile % n_tiles) * FU_BN;
      {
        const int ch = n0 + s * 64 + 2 * lane;
        float2 w[9];
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) w[tp] = __ldg(reinterpret_cast<const float2*>(ep.dw_w9 + (size_t)tp * N + ch));
        const float2 bs = __ldg(reinterpret_cast<const float2*>(ep.dw_b + ch));
        const uint32_t rows = smem_u32(smem_tile) + s * FU_TILE_BYTES + (2 * j) * 16 * 128;
        mbar_wait(&cfull_bar[s], uint32_t(it) & 1u);
        auto ldcol = [&](float2 (&dst)[4], int x) {
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            uint32_t v;
            const uint32_t addr = rows + (rr * 16 + x) * 128 + ((lane_chunk ^ (x & 7)) << 4) + lane_off;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
            dst[rr] = unpack_bf16x2(v);
          }
        };
        float2 win[3][4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) win[2][rr] = make_float2(0.f, 0.f);  // column -1
        ldcol(win[0], 0);
        ldcol(win[1], 1);
        bf16* out = ep.out + ((size_t)m0 + (size_t)(2 * j) * 16) * N + ch;
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          float2 a0 = bs, a1 = bs;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              a0 = ffma2(w[dy * 3 + dx], win[(x + dx + 2) % 3][dy], a0);
              a1 = ffma2(w[dy * 3 + dx], win[(x + dx + 2) % 3][dy + 1], a1);
            }
          const float2 g0 = gelu2_logistic(a0), g1 = gelu2_logistic(a1);   // MUFU ex2 + rcp: the FMA pipe is this kernel's bound
          *reinterpret_cast<uint32_t*>(out + (size_t)x * N) = pack_bf16x2(g0.x, g0.y);
          *reinterpret_cast<uint32_t*>(out + (size_t)(16 + x) * N) = pack_bf16x2(g1.x, g1.y);
          if (x + 2 < 16) {
            ldcol(win[(x + 2) % 3], x + 2);
          } else {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) win[(x + 2) % 3][rr] = make_float2(0.f, 0.f);
          }
        }
        __syncwarp();
        if (lane == 0) {   // tile buffer consumed: release it to our drain warps and to the peer's boundary warp
          mbar_arrive(&cempty_bar[s]);
          mbar_arrive_cluster(mapa_u32(smem_u32(&cempty_bar[s]), peer));
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // nobody exits while the peer may still push a halo row or arrive on one of our barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// g[M, N] = GELU(dwconv3x3(A[M,K] W[N,K]^T (LN-folded) + c) + dw_b), M = batch * 256 tokens (16 x 16 grid per sample)
int launch_gemm_up_dwconv_gelu(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const float* col_c,
                               const float* col_s, const float2* row_part, int n_part, float ln_eps, const float* dw_w9,
                               const float* dw_b, bf16* out, cudaStream_t st, bf16* hid_out) {
  TLD_CHECK(M > 0 && M % 256 == 0, "gemm_up_dwconv: rows must be whole 16x16-token samples (M % 256 == 0)");
  TLD_CHECK(N > 0 && N % FU_BN == 0, "gemm_up_dwconv: hidden width must be a multiple of 256");
  TLD_CHECK(K > 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm_up_dwconv: K/lda/ldw must be multiples of 8");
  TLD_CHECK(col_c && dw_w9 && dw_b && out, "gemm_up_dwconv: null argument");
  TLD_CHECK((col_s == nullptr) == (row_part == nullptr) && (row_part == nullptr || (n_part > 0 && n_part <= 32)),
            "gemm_up_dwconv: the LayerNorm fold needs both col_s and the row partials");
  TLD_CHECK(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(out) |
              reinterpret_cast<uintptr_t>(col_c) | reinterpret_cast<uintptr_t>(col_s) | reinterpret_cast<uintptr_t>(dw_w9) |
              reinterpret_cast<uintptr_t>(dw_b)) & 15) == 0 && (reinterpret_cast<uintptr_t>(row_part) & 7) == 0,
            "gemm_up_dwconv: operands must be 16-byte aligned");
  CUtensorMap ta, tb;
  if (make_tmap_2d(&ta, A, false, M, K, lda, 128)) return 1;
  if (make_tmap_2d(&tb, W, false, N, K, ldw, FU_BN / 2)) return 1;
  FusedUpArgs ep{};
  ep.col_c = col_c;
  ep.col_s = col_s;
  ep.row_part = row_part;
  ep.n_part = n_part;
  ep.inv_d = 1.f / float(K);
  ep.ln_eps = ln_eps;
  ep.dw_w9 = dw_w9;
  ep.dw_b = dw_b;
  ep.out = out;
  ep.hid_out = hid_out;
  auto kern = gemm_up_dwconv_gelu_kernel;
  static bool attr_set = false;
  if (!attr_set) {
    TLD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_SMEM));
    attr_set = true;
  }
  const int tiles = (M / 256) * (N / FU_BN);
  const int slots = sm_count() / 2;
  const int grid = (tiles < slots ? tiles : slots) * 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(FU_THREADS);
  cfg.dynamicSmemBytes = FU_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  TLD_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, M, N, K, ep));
  return 0;
}

}  // namespace tld
