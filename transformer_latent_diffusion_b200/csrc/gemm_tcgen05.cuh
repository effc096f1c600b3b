// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
//   A : bf16 row-major [M,K] (activations, K contiguous)       -> TMA, 128B-swizzled smem tiles 128x64
//   W : bf16 row-major [N,K] (nn.Linear / 1x1-conv weight)     -> TMA, 128B-swizzled smem tiles BNx64
//   accumulators: fp32 in TMEM, two stages of BN columns so the epilogue of tile i overlaps the MMAs of tile i+1
//   C : written by TMA from 128B-swizzled shared-memory staging: plain tiled store for fresh outputs,
//       TMA reduce-add (read-modify-write at L2, no x read through the SM) for the residual epilogues
//
// Warp roles (256 threads, 1 CTA per SM, grid = min(#tiles, #SMs), static round-robin tile schedule):
//   warp 0      TMA producer (one elected lane)
//   warp 1      MMA issuer   (one elected lane issues tcgen05.mma; tcgen05.commit releases smem / publishes TMEM)
//   warp 2      TMEM allocator / deallocator
//   warps 4..7  epilogue: each warp owns 32 accumulator rows (its TMEM lane quarter): tcgen05.ld -> fused math
//               -> its own [32 rows x 128 B] staging slab (double buffered) -> its own TMA store / reduce-add.
//               No cross-warp synchronisation on the store path.
//
// CTAS = 2 (cta_group::2): two CTAs of a cluster (an SM pair) share one 256 x BN tile.  Each CTA stages its own
// 128 A rows and HALF of the W tile (BN/2 rows); the leader CTA issues tcgen05.mma.cta_group::2 with M = 256, each
// SM's tensor core reads W halves from both SMs' shared memory, and every stage moves 32 KB instead of 48 KB per SM
// at BN = 256 -- the single-CTA kernel is shared-memory-bandwidth bound (~65-70 % tensor-pipe active in ncu).
// Barriers: TMA of both CTAs completes on the leader's full barrier; tcgen05.commit multicasts to both CTAs'
// empty / tmem-full barriers; both CTAs' epilogue warps arrive on the leader's tmem-empty barrier.
//
// Epilogues (reference call sites, /root/reference/tld/transformer_blocks.py):
//   EPI_BF16            out_bf16 = acc                                   qkv_linear            (:58)
//   EPI_BIAS_BF16       out_bf16 = acc + bias                            mlp.0 1x1 conv D->4D  (:95)
//   EPI_BIAS_RESID_F32  x_f32   += acc + bias                            mlp.3 1x1 conv + "+x" (:104,:138)
//   EPI_XATTN_RESID_F32 x_f32   += softmax2(q.k0, q.k1) . (v0, v1)       q_linear + 2-key SDPA + "+x" (:70-72,:137)
//   EPI_F32             out_f32  = acc                                   kv_linear on cond tokens (:71)
// LayerNorm folding (the norm1 / norm3 kernels disappear; their statistics ride on the residual epilogues):
//   EPI_LNFOLD_BF16     out_bf16 = rstd_r (acc - mean_r s_n) + c_n               consumer: A = bf16(x) un-normalised, W = gamma (.) W,
//                        s_n = sum_k W'_nk, c_n = sum_k beta_k W_nk (+ bias); mean / rstd from per-32-column (sum, sum of
//                        squares) partials of the fp32 row written by the producer of x
//   EPI_BIAS_RESID_LNP  x_f32 = x_f32 + acc + bias (explicit read-modify-write), plus the bf16 copy of the new row (tmap_d) and its
//   EPI_XATTN_RESID_LNP per-32-column (sum, sum of squares) partials: the producer side, for mlp.3 and for the cross-attention
#pragma once
#include "ptx.cuh"

namespace tld {

enum EpiMode : int {
  EPI_BF16 = 0,
  EPI_BIAS_BF16 = 1,
  EPI_BIAS_RESID_F32 = 2,
  EPI_XATTN_RESID_F32 = 3,
  EPI_F32 = 4,
  EPI_LNFOLD_BF16 = 5,
  EPI_BIAS_RESID_LNP = 6,
  EPI_XATTN_RESID_LNP = 7,
};

struct GemmEpi {
  const float* bias;  // [N] or nullptr
  // cross-attention epilogue only
  const float* kv0;   // cond token 0 (noise) K|V rows for this layer: K at [0,D), V at [D,2D)
  const float* kv1;   // cond token 1 (label)
  long long kv0_stride, kv1_stride;  // floats between consecutive rows
  const int* step_ptr;  // if non-null: every sample uses kv0 row *step_ptr (hoisted per-step noise token)
  int n_tok;          // tokens per sample (row -> sample = row / n_tok)
  int embed_dim;      // D
  float scale;        // 1/sqrt(head_dim)
  // implicit-GEMM 3x3 convolution (conv_cpb > 0): A rows are NHWC pixels, K runs over (tap, 64-channel block);
  // the A tile of a k-block is a 4-D TMA box [64 ch, w_box, h_box, 1] shifted by the tap, zero-filled outside.
  int conv_cpb;       // Cin / 64 (0 = plain GEMM)
  int conv_h, conv_w; // image height / width in pixels
  // mn_major bit 0: A is stored [K, M], bit 1: B is stored [K, N] (row-major over K).  3 = the wgrad shape dW = dY^T X with
  // K = tokens; 2 = the dgrad shape dX = dY W with the weight W [N_out = K, N_in = N] as it is.  Such an operand is loaded as
  // [64 k x 64 mn] TMA boxes and fed to tcgen05.mma as an MN-major tile (idesc a/b_major = 1), so no transposed copies of
  // activations or weights are ever written.
  int mn_major;
  // LayerNorm fold, consumer side (EPI_LNFOLD_BF16): bias = c_n, col_s = s_n, row_part [M, n_part] partial row statistics
  const float* col_s;
  const float2* row_part;
  int n_part;               // K / 32 partials per row
  float inv_d, ln_eps;      // 1 / K and the LayerNorm epsilon
  // producer side (EPI_*_LNP): x_in == the fp32 output (read thread = row before it is overwritten), part_out [M, N / 32]
  const float* x_in;
  float2* part_out;
  // EPI_BIAS_BF16 extras (VAE convolutions): resid = bf16 [M, N] added before the rounding (ResnetBlock shortcut); gn_part =
  // [M / 32, N / 4] (sum, sum of squares) of the STORED bf16 values per 32-row slab and 4-channel quad: the GroupNorm that
  // consumes this tensor sums them instead of reading it once more
  const __nv_bfloat16* resid;
  float* gn_part;
};

// 32 per-lane values -> lane L ends with the warp total of value L (16 + 8 + 4 + 2 + 1 shuffles)
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool hi = lane & w;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float keep = hi ? v[w + i] : v[i], send = hi ? v[i] : v[w + i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return v[0];
}

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_THREADS = 256;
constexpr int UMMA_K = 16;
constexpr int STG_SLAB = 32 * 128;  // one epilogue warp's staging slab: 32 rows x 128 B

template <int BN, int EPI, int CTAS = 1>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = (BN / CTAS) * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr bool LNP = (EPI == EPI_BIAS_RESID_LNP || EPI == EPI_XATTN_RESID_LNP);
  // plain modes: 4 warps x 2 staging slabs.  Producer modes: 4 warps x (LNP_SLABS in-place fp32 slabs: TMA load of x_old ->
  // read-modify-write in shared memory -> TMA store of x_new, 2 chunks of prefetch) + 2 slabs for the bf16 copy
  // mlp.3 (K = 4D, long main loop, epilogue has slack): 2 in-place slabs + 1 bf16 slab keep 6 pipeline stages for the
  // DRAM-streamed A operand; cross-attention (K = D, epilogue-bound): 4 in-place slabs (2 chunks of prefetch) + 2 bf16 slabs
  static constexpr int LNP_SLABS = EPI == EPI_XATTN_RESID_LNP ? 4 : 2;
  static constexpr int LNP_BF16_SLABS = EPI == EPI_XATTN_RESID_LNP ? 2 : 1;
  static constexpr int STG_BYTES = LNP ? 4 * (LNP_SLABS + LNP_BF16_SLABS) * STG_SLAB : 4 * 2 * STG_SLAB;
  // cross-attention: 2 buffers x 2 samples x {k0,k1,v0,v1};  LayerNorm-fold consumer: 2 buffers x {c_n, s_n} of the tile's columns
  static constexpr int KV_BYTES = (EPI == EPI_XATTN_RESID_F32 || EPI == EPI_XATTN_RESID_LNP) ? 2 * 2 * 4 * BN * 4
                                  : (EPI == EPI_LNFOLD_BF16 ? 2 * 2 * BN * 4 : 0);
  static constexpr int BUDGET = 227 * 1024 - 1024 /*align slack*/ - STG_BYTES - KV_BYTES - 512 /*barriers*/;
  static constexpr int STAGES = (BUDGET / STAGE_BYTES) > 8 ? 8 : (BUDGET / STAGE_BYTES);
  static constexpr int TOTAL = 1024 + STAGES * STAGE_BYTES + STG_BYTES + KV_BYTES + 512;
  // producer modes keep 96 KB of slabs: their wide single-CTA tiles would be left with 2 stages and are never dispatched
  // (launch_gemm narrows BN to <= 128 for them)
  static_assert(STAGES >= (LNP ? 2 : 3), "not enough shared memory for the pipeline");
};

// write this lane's 128-byte row into a 128B-swizzled [32 x 128 B] slab (conflict-free per quarter-warp)
__device__ __forceinline__ void stage_row(uint8_t* slab, int lane, const uint32_t (&v)[32]) {
  const uint32_t row = smem_u32(slab) + lane * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j) sts_v4(row + ((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

template <int BN, int EPI, int CTAS>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_d, int M, int N, int K,
                    GemmEpi ep) {
  using S = GemmSmem<BN, EPI, CTAS>;
  static_assert(CTAS == 1 || CTAS == 2, "CTAS must be 1 or 2");
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  constexpr int STAGES = S::STAGES;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                 : (2 * BN <= 256) ? 256 : 512;
  constexpr bool OUT_BF16 = (EPI == EPI_BF16 || EPI == EPI_BIAS_BF16 || EPI == EPI_LNFOLD_BF16);
  constexpr bool REDUCE = (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_XATTN_RESID_F32);
  constexpr bool XATTN = (EPI == EPI_XATTN_RESID_F32 || EPI == EPI_XATTN_RESID_LNP);
  constexpr bool LNP = (EPI == EPI_BIAS_RESID_LNP || EPI == EPI_XATTN_RESID_LNP);
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "BN must be 64..256, multiple of 64");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * S::A_BYTES;
  uint8_t* smem_stg = smem + STAGES * S::STAGE_BYTES;  // 1024-aligned: every stage size is a multiple of 1024
  float* smem_kv = reinterpret_cast<float*>(smem_stg + S::STG_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + S::STG_BYTES + S::KV_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* lnp_bar = bars + 2 * STAGES + 5;    // [4 warps][LNP_SLABS]: x_old slab loaded (producer modes)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int TILE_M = GEMM_BM * CTAS;  // rows per (cluster) tile
  const int m_tiles = (M + TILE_M - 1) / TILE_M;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + GEMM_BK - 1) / GEMM_BK;
  const int first_tile = blockIdx.x / CTAS;      // both CTAs of a pair walk the same tile sequence
  const int tile_step = gridDim.x / CTAS;

  pdl_launch_dependents();
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
    if constexpr (LNP) tma_prefetch_desc(&tmap_d);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);      // the leader's expect_tx arrive; the peer's TMA only adds complete_tx bytes
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4 * CTAS);  // one arrive per epilogue warp of every CTA in the pair
    }
    if constexpr (LNP)
      for (int s = 0; s < 4 * S::LNP_SLABS; ++s) mbar_init(&lnp_bar[s], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (CTAS == 2) {
      tmem_alloc_pair(tmem_slot, TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  pdl_wait();  // prologue above overlaps the previous kernel's tail; operands / outputs are touched only from here on
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int m0 = (tile / n_tiles) * TILE_M + int(cta_rank) * GEMM_BM;
        const int n0 = (tile % n_tiles) * BN + int(cta_rank) * (BN / CTAS);
        if constexpr (LNP) {
          // producer modes: pull the fp32 residual tile the epilogue will read-modify-write one tile from now into L2, so
          // that its slab loads see L2 latency instead of DRAM latency
          const int tn = tile + tile_step;
          if (tn < num_tiles) {
            const int pm = (tn / n_tiles) * TILE_M + int(cta_rank) * GEMM_BM, pn = (tn % n_tiles) * BN;
            for (int r = 0; r < GEMM_BM; r += 32)
              for (int c = 0; c < BN; c += 32)
                if (pm + r < M && pn + c < N) tma_prefetch_2d(&tmap_c, pn + c, pm + r);
          }
        }
        // implicit-GEMM conv: first pixel of this CTA's 128-row tile -> (image, y, x)
        int cimg = 0, cy0 = 0, cx0 = 0;
        if (ep.conv_cpb > 0) {
          const int hw = ep.conv_h * ep.conv_w;
          cimg = m0 / hw;
          const int rem = m0 - cimg * hw;
          cy0 = rem / ep.conv_w;
          cx0 = rem - cy0 * ep.conv_w;
        }
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          int tap = 0, cb = kb;
          if (ep.conv_cpb > 0) {
            tap = kb / ep.conv_cpb;
            cb = kb - tap * ep.conv_cpb;
          }
          const int ky = tap / 3, kx = tap - 3 * ky;
          if (ep.mn_major) {
            // MN-major operand (bit 0: A, bit 1: B): one [64 k-rows x 64 elements] box (8 KB, 128B-swizzled rows) per 64
            // columns of the tile; the other operand, if K-major, takes the usual single box
            const bool a_mn = ep.mn_major & 1, b_mn = ep.mn_major & 2;
            if constexpr (CTAS == 2) {
              if (leader) mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES * 2);
              if (a_mn) {
#pragma unroll
                for (int j = 0; j < GEMM_BM / 64; ++j)
                  tma_load_2d_pair(smem_a + stage * S::A_BYTES + j * 8192, &tmap_a, &full_bar[stage], m0 + 64 * j, kb * GEMM_BK);
              } else {
                tma_load_2d_pair(smem_a + stage * S::A_BYTES, &tmap_a, &full_bar[stage], kb * GEMM_BK, m0);
              }
              if (b_mn) {
#pragma unroll
                for (int j = 0; j < BN / CTAS / 64; ++j)
                  tma_load_2d_pair(smem_b + stage * S::B_BYTES + j * 8192, &tmap_b, &full_bar[stage], n0 + 64 * j, kb * GEMM_BK);
              } else {
                tma_load_2d_pair(smem_b + stage * S::B_BYTES, &tmap_b, &full_bar[stage], kb * GEMM_BK, n0);
              }
            } else {
              mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
              if (a_mn) {
#pragma unroll
                for (int j = 0; j < GEMM_BM / 64; ++j)
                  tma_load_2d(smem_a + stage * S::A_BYTES + j * 8192, &tmap_a, &full_bar[stage], m0 + 64 * j, kb * GEMM_BK);
              } else {
                tma_load_2d(smem_a + stage * S::A_BYTES, &tmap_a, &full_bar[stage], kb * GEMM_BK, m0);
              }
              if (b_mn) {
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                  tma_load_2d(smem_b + stage * S::B_BYTES + j * 8192, &tmap_b, &full_bar[stage], n0 + 64 * j, kb * GEMM_BK);
              } else {
                tma_load_2d(smem_b + stage * S::B_BYTES, &tmap_b, &full_bar[stage], kb * GEMM_BK, n0);
              }
            }
          } else if constexpr (CTAS == 2) {
            // Only the leader arms its barrier (for both CTAs' bytes).  The peer may run at most one phase ahead:
            // its own empty barrier is released by the commit that follows the MMAs of the previous phase.
            if (leader) mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES * 2);
            if (ep.conv_cpb > 0)
              tma_load_4d_pair(smem_a + stage * S::A_BYTES, &tmap_a, &full_bar[stage], cb * GEMM_BK, cx0 + kx - 1,
                               cy0 + ky - 1, cimg);
            else
              tma_load_2d_pair(smem_a + stage * S::A_BYTES, &tmap_a, &full_bar[stage], kb * GEMM_BK, m0);
            tma_load_2d_pair(smem_b + stage * S::B_BYTES, &tmap_b, &full_bar[stage], kb * GEMM_BK, n0);
          } else {
            mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
            if (ep.conv_cpb > 0)
              tma_load_4d(smem_a + stage * S::A_BYTES, &tmap_a, &full_bar[stage], cb * GEMM_BK, cx0 + kx - 1,
                          cy0 + ky - 1, cimg);
            else
              tma_load_2d(smem_a + stage * S::A_BYTES, &tmap_a, &full_bar[stage], kb * GEMM_BK, m0);
            tma_load_2d(smem_b + stage * S::B_BYTES, &tmap_b, &full_bar[stage], kb * GEMM_BK, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader CTA only when paired) =====================
    const bool a_mn = ep.mn_major & 1, b_mn = ep.mn_major & 2;
    const uint32_t idesc = a_mn ? (b_mn ? umma_idesc_bf16(TILE_M, BN, 1, 1) : umma_idesc_bf16(TILE_M, BN, 1, 0))
                                : (b_mn ? umma_idesc_bf16(TILE_M, BN, 0, 1) : umma_idesc_bf16(TILE_M, BN, 0, 0));
    // K-major tiles: 8-row groups 1024 B apart, +32 B per K=16 step inside the swizzle row.  MN-major tiles: 64-wide
    // column atoms 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO), +16 k-rows = 2048 B per K=16 step.
    const uint32_t lbo_a = a_mn ? 8192u : 16u, kstep_a = a_mn ? 128u : 2u;
    const uint32_t lbo_b = b_mn ? 8192u : 16u, kstep_b = b_mn ? 128u : 2u;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = umma_smem_desc_sw128(smem_u32(smem_a + stage * S::A_BYTES), lbo_a, 1024);
          const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(smem_b + stage * S::B_BYTES), lbo_b, 1024);
#pragma unroll
          for (int k = 0; k < GEMM_BK / UMMA_K; ++k) {
            // advance 32 B (16 bf16) along K inside the 128 B swizzle row: +2 in the (addr>>4) field
            if constexpr (CTAS == 2) umma_ss_f16_pair(d_tmem, adesc + kstep_a * k, bdesc + kstep_b * k, idesc, (kb | k) != 0);
            else umma_ss_f16(d_tmem, adesc + kstep_a * k, bdesc + kstep_b * k, idesc, (kb | k) != 0);
          }
          if constexpr (CTAS == 2) {
            umma_commit_pair(&empty_bar[stage]);                      // frees this stage in BOTH CTAs
            if (kb == k_blocks - 1) umma_commit_pair(&tfull_bar[acc]);  // accumulator halves complete in both CTAs
          } else {
            umma_commit(&empty_bar[stage]);                      // smem slot free once these MMAs retire
            if (kb == k_blocks - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;                 // == warp % 4: the TMEM lane quarter this warp may read
    const int et = threadIdx.x - 128;        // 0..127 == accumulator row inside the tile
    uint8_t* my_stg = smem_stg + ew * 2 * STG_SLAB;
    int buf = 0;
    // push one staged [32 rows x 128 B] slab to global: plain store or reduce-add; ordering: lane 0 makes sure the
    // slab it is about to hand out again has been read by the TMA unit before anyone overwrites it.
    auto slab_acquire = [&]() -> uint8_t* {
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
      return my_stg + buf * STG_SLAB;
    };
    auto slab_release = [&](uint8_t* slab, int col, int row) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && row < M && col < N) {  // fully out-of-range slabs are skipped, partial ones are clipped by TMA
        if constexpr (REDUCE) tma_reduce_add_2d(&tmap_c, slab, col, row);
        else tma_store_2d(&tmap_c, slab, col, row);
        bulk_commit();
      }
      buf ^= 1;
    };
    // ---- producer modes: the fp32 residual goes through shared memory in place.  Per warp LNP_SLABS slabs [32 rows x 32
    // columns fp32, 128B-swizzled]: chunk q of this warp's chunk sequence (tiles in order, BN/32 chunks each) uses slab q % 4;
    // x_old arrives by TMA (issued two chunks ahead), every lane adds its row's new values in place, the same slab leaves by
    // TMA store.  One bulk group per chunk and one per 64-column bf16 copy (committed even when nothing is stored), so that
    // "all but the 2 latest groups have been read" == the slab of chunk q - 2 is free to be refilled with chunk q + 2.
    constexpr int NCH = BN / 32;
    uint8_t* pslab = smem_stg + ew * (S::LNP_SLABS + S::LNP_BF16_SLABS) * STG_SLAB;
    uint64_t* my_ld_bar = lnp_bar + ew * S::LNP_SLABS;
    int q_chunk = 0;   // chunks this warp has processed so far
    int q_pair = 0;    // 64-column bf16 copies this warp has stored so far (their two slabs alternate strictly)
    auto lnp_issue = [&](int tile_, int ch_, int q_) {   // lane 0: fetch x_old of chunk (tile_, ch_) into slab q_ % LNP_SLABS
      const int srow = (tile_ / n_tiles) * TILE_M + int(cta_rank) * GEMM_BM + ew * 32;
      const int col = (tile_ % n_tiles) * BN + ch_ * 32;
      if (srow < M && col < N) {
        const int sl = q_ % S::LNP_SLABS;
        mbar_expect_tx(&my_ld_bar[sl], STG_SLAB);
        tma_load_2d(pslab + sl * STG_SLAB, &tmap_c, &my_ld_bar[sl], col, srow);
      }
    };
    // x_old of this lane's row, chunk q_chunk (zeros if the slab is out of range); the caller overwrites it with x_new
    auto lnp_wait = [&](int srow, int col) -> uint32_t {
      const int sl = q_chunk % S::LNP_SLABS;
      if (srow < M && col < N) mbar_wait(&my_ld_bar[sl], uint32_t(q_chunk / S::LNP_SLABS) & 1u);
      return smem_u32(pslab + sl * STG_SLAB) + lane * 128;
    };
    // start of a chunk: with 4 slabs recycle the slab of chunk q - 2 for chunk q + 2 (at most 2 younger groups exist: no
    // stall); with 2 slabs the other slab (chunk q - 1, every group read) is refilled with chunk q + 1
    auto lnp_advance = [&](int tile_, int ch_) {
      constexpr int AHEAD = S::LNP_SLABS == 4 ? 2 : 1;
      if (lane == 0) {
        if constexpr (S::LNP_SLABS == 4) bulk_wait_read<2>();
        else bulk_wait_read<0>();
        int ch2 = ch_ + AHEAD, tile2 = tile_;
        if (ch2 >= NCH) { ch2 -= NCH; tile2 += tile_step; }
        if (tile2 < num_tiles) lnp_issue(tile2, ch2, q_chunk + AHEAD);
      }
      __syncwarp();
    };
    auto lnp_store = [&](uint32_t slab_lane_addr, int srow, int col) {   // after the lanes rewrote their rows in place
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (srow < M && col < N) tma_store_2d(&tmap_c, pslab + (q_chunk % S::LNP_SLABS) * STG_SLAB, col, srow);
        bulk_commit();
      }
      ++q_chunk;
    };
    auto lnp_store_bf16 = [&](const uint32_t (&ob)[32], int srow, int col) {
      // previous use of this slab: 6 groups back with two slabs, 3 groups back (and every group read) with one
      uint8_t* slab = pslab + (S::LNP_SLABS + (q_pair % S::LNP_BF16_SLABS)) * STG_SLAB;
      ++q_pair;
      stage_row(slab, lane, ob);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (srow < M && col < N) tma_store_2d(&tmap_d, slab, col, srow);
        bulk_commit();
      }
    };
    if constexpr (LNP) {
      if (lane == 0 && first_tile < num_tiles) {
        lnp_issue(first_tile, 0, 0);
        if constexpr (S::LNP_SLABS == 4) lnp_issue(first_tile, 1, 1);
      }
    }

    // cross-attention K/V staging (EPI_XATTN_RESID_F32 only): 2 samples x 4 vectors x BN columns as float4 pieces
    constexpr int KV_LD = (2 * BN + 127) / 128;   // float4 loads per thread
    float4 kv_pre[KV_LD];
    const long long r0s = (XATTN && ep.step_ptr) ? (long long)(*ep.step_ptr) : -1;
    auto kv_fetch = [&](int tile) {
      const int m0 = (tile / n_tiles) * TILE_M + int(cta_rank) * GEMM_BM;
      const int n0 = (tile % n_tiles) * BN;
      const int b_first = m0 / ep.n_tok;
#pragma unroll
      for (int u = 0; u < KV_LD; ++u) {
        const int j = et + 128 * u;
        const int s = j / BN, rem = j % BN;          // BN float4 per sample
        const int vec = rem / (BN / 4), c = (rem % (BN / 4)) * 4;   // 0:k0 1:k1 2:v0 3:v1
        const int b = b_first + s;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < 2 * BN && (long long)b * ep.n_tok < M && n0 + c < N) {
          const long long r0 = r0s >= 0 ? r0s : (long long)b;
          const float* src = (vec & 1) == 0 ? ep.kv0 + r0 * ep.kv0_stride : ep.kv1 + (long long)b * ep.kv1_stride;
          val = __ldg(reinterpret_cast<const float4*>(src + (vec >= 2 ? ep.embed_dim : 0) + n0 + c));
        }
        kv_pre[u] = val;
      }
    };
    auto kv_store = [&](uint32_t base) {
#pragma unroll
      for (int u = 0; u < KV_LD; ++u) {
        const int j = et + 128 * u;
        if (j < 2 * BN)
          sts_v4(base + j * 16, __float_as_uint(kv_pre[u].x), __float_as_uint(kv_pre[u].y), __float_as_uint(kv_pre[u].z),
                 __float_as_uint(kv_pre[u].w));
      }
    };
    if constexpr (XATTN) {
      if (first_tile < num_tiles) kv_fetch(first_tile);
    }

    int it = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = (tile / n_tiles) * TILE_M + int(cta_rank) * GEMM_BM;   // this CTA's 128 rows
      const int n0 = (tile % n_tiles) * BN;
      const int row = m0 + et;
      const int slab_row = m0 + ew * 32;

      int sidx = 0;
      uint32_t kvb = 0;
      if constexpr (XATTN) {
        // k0,k1,v0,v1 of the (at most two) samples this tile touches: fetched into registers one tile ahead (the global
        // loads are in flight during the previous tile's epilogue math), parked in smem buffer it & 1.  One barrier per
        // tile suffices: whoever writes buffer it & 1 has passed barrier it-1, which every reader of tile it-2 reached
        // only after its reads.
        kvb = smem_u32(smem_kv) + (it & 1) * (2 * 4 * BN * 4);
        kv_store(kvb);
        named_bar_sync(1, 128);
        if (tile + tile_step < num_tiles) kv_fetch(tile + tile_step);
        sidx = row < M ? (row / ep.n_tok - m0 / ep.n_tok) : 0;
        kvb += sidx * 4 * BN * 4;
      }

      float ln_a = 1.f, ln_b = 0.f;   // EPI_LNFOLD_BF16: value = ln_a * acc + (ln_b * s_n + c_n)
      uint32_t cs_u32 = 0;
      if constexpr (EPI == EPI_LNFOLD_BF16) {
        // the tile's column constants through shared memory (every thread needs every column: broadcast LDS.128 instead of
        // two uniform global loads per column pair).  Buffer it & 1; the barrier of tile it + 1 fences its readers.
        float* cs = smem_kv + (it & 1) * (2 * BN);
#pragma unroll
        for (int u = 0; u < (BN + 127) / 128; ++u) {
          const int c = et + 128 * u;
          if (c < BN) {
            const bool ok = n0 + c < N;
            cs[c] = ok ? __ldg(ep.bias + n0 + c) : 0.f;
            cs[BN + c] = ok ? __ldg(ep.col_s + n0 + c) : 0.f;
          }
        }
        named_bar_sync(1, 128);
        cs_u32 = smem_u32(cs);
        if (row < M) {
          const float4* pp = reinterpret_cast<const float4*>(ep.row_part + (size_t)row * ep.n_part);
          float4 pv[16];   // all partials of the row in flight at once (K <= 1024): a dependent load chain per tile costs more
#pragma unroll       // than the tile's MMAs
          for (int u = 0; u < 16; ++u) pv[u] = u < ep.n_part / 2 ? __ldg(pp + u) : make_float4(0.f, 0.f, 0.f, 0.f);
          float ps = 0.f, pq = 0.f;
#pragma unroll
          for (int u = 0; u < 16; ++u) {   // fixed order: deterministic
            ps += pv[u].x + pv[u].z;
            pq += pv[u].y + pv[u].w;
          }
          const float mean = ps * ep.inv_d;
          ln_a = rsqrtf(fmaxf(pq * ep.inv_d - mean * mean, 0.f) + ep.ln_eps);
          ln_b = -ln_a * mean;
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + acc * BN;

      if constexpr (XATTN) {
#pragma unroll 1
        for (int hc = 0; hc < BN / 64; ++hc) {
          uint32_t qa[32], qb[32];
          tmem_ld_x32(taddr + hc * 64, qa);
          tmem_ld_x32(taddr + hc * 64 + 32, qb);
          tmem_ld_wait();
          const uint32_t k0 = kvb + (0 * BN + hc * 64) * 4, k1 = kvb + (1 * BN + hc * 64) * 4;
          const uint32_t v0 = kvb + (2 * BN + hc * 64) * 4, v1 = kvb + (3 * BN + hc * 64) * 4;
          float s0a = 0.f, s0b = 0.f, s1a = 0.f, s1b = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 a = lds_v4(k0 + i * 16), b = lds_v4(k1 + i * 16), c = lds_v4(k0 + (8 + i) * 16), d = lds_v4(k1 + (8 + i) * 16);
            const float x0 = __uint_as_float(qa[4 * i]), x1 = __uint_as_float(qa[4 * i + 1]);
            const float x2 = __uint_as_float(qa[4 * i + 2]), x3 = __uint_as_float(qa[4 * i + 3]);
            const float y0 = __uint_as_float(qb[4 * i]), y1 = __uint_as_float(qb[4 * i + 1]);
            const float y2 = __uint_as_float(qb[4 * i + 2]), y3 = __uint_as_float(qb[4 * i + 3]);
            s0a += x0 * a.x + x1 * a.y + x2 * a.z + x3 * a.w;
            s1a += x0 * b.x + x1 * b.y + x2 * b.z + x3 * b.w;
            s0b += y0 * c.x + y1 * c.y + y2 * c.z + y3 * c.w;
            s1b += y0 * d.x + y1 * d.y + y2 * d.z + y3 * d.w;
          }
          const float s0 = (s0a + s0b) * ep.scale, s1 = (s1a + s1b) * ep.scale;
          const float mx = fmaxf(s0, s1);
          const float e0 = __expf(s0 - mx), e1 = __expf(s1 - mx);
          const float inv = 1.f / (e0 + e1);
          const float p0 = e0 * inv, p1 = e1 * inv;
          uint32_t ob[LNP ? 32 : 1];   // producer mode: the 64 columns of this head as bf16 pairs
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int col = n0 + hc * 64 + half * 32;
            uint32_t xin = 0;
            if constexpr (LNP) {
              lnp_advance(tile, hc * 2 + half);
              xin = lnp_wait(slab_row, col);
            }
            const bool have_x = LNP && slab_row < M && col < N;
            uint32_t o[32];
            float ps = 0.f, pq = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 a = lds_v4(v0 + (half * 8 + i) * 16), b = lds_v4(v1 + (half * 8 + i) * 16);
              float w0 = p0 * a.x + p1 * b.x, w1 = p0 * a.y + p1 * b.y, w2 = p0 * a.z + p1 * b.z, w3 = p0 * a.w + p1 * b.w;
              if constexpr (LNP) {
                if (have_x) {
                  const float4 xo = lds_v4(xin + ((i ^ (lane & 7)) << 4));
                  w0 += xo.x; w1 += xo.y; w2 += xo.z; w3 += xo.w;
                }
                ps += (w0 + w1) + (w2 + w3);
                pq += (w0 * w0 + w1 * w1) + (w2 * w2 + w3 * w3);
                ob[half * 16 + 2 * i] = pack_bf16x2(w0, w1);
                ob[half * 16 + 2 * i + 1] = pack_bf16x2(w2, w3);
              }
              o[4 * i] = __float_as_uint(w0);
              o[4 * i + 1] = __float_as_uint(w1);
              o[4 * i + 2] = __float_as_uint(w2);
              o[4 * i + 3] = __float_as_uint(w3);
            }
            if constexpr (LNP) {
              if (row < M && col < N) ep.part_out[(size_t)row * (N / 32) + col / 32] = make_float2(ps, pq);
#pragma unroll
              for (int j = 0; j < 8; ++j) sts_v4(xin + ((j ^ (lane & 7)) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
              lnp_store(xin, slab_row, col);
            } else {
              uint8_t* slab = slab_acquire();
              stage_row(slab, lane, o);
              slab_release(slab, col, slab_row);
            }
          }
          if constexpr (LNP) lnp_store_bf16(ob, slab_row, n0 + hc * 64);
        }
      } else if constexpr (OUT_BF16) {
#pragma unroll 1
        for (int ch = 0; ch < BN / 64; ++ch) {
          uint32_t ra[32], rb[32];
          tmem_ld_x32(taddr + ch * 64, ra);
          tmem_ld_x32(taddr + ch * 64 + 32, rb);
          tmem_ld_wait();
          const int col = n0 + ch * 64;
          uint32_t o[32];
          [[maybe_unused]] uint32_t rw[32];   // residual: the 64 bf16 of this row under the chunk
          [[maybe_unused]] bool has_res = false;
          if constexpr (EPI == EPI_BIAS_BF16) {
            has_res = ep.resid != nullptr && row < M && col + 63 < N;
            if (has_res) {
              const uint4* rp = reinterpret_cast<const uint4*>(ep.resid + (size_t)row * N + col);
#pragma unroll
              for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(&rw[4 * j]) = __ldg(rp + j);
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float f0 = __uint_as_float(ra[2 * i]), f1 = __uint_as_float(ra[2 * i + 1]);
            float g0 = __uint_as_float(rb[2 * i]), g1 = __uint_as_float(rb[2 * i + 1]);
            if constexpr (EPI == EPI_BIAS_BF16) {
              if (has_res) {
                f0 += __uint_as_float(rw[i] << 16); f1 += __uint_as_float(rw[i] & 0xffff0000u);
                g0 += __uint_as_float(rw[16 + i] << 16); g1 += __uint_as_float(rw[16 + i] & 0xffff0000u);
              }
              // N is a multiple of 32: a ragged last group (N % 64 == 32) still gets its first half's bias
              if (col + 31 < N) {
                const float2 b0 = __ldg(reinterpret_cast<const float2*>(ep.bias + col) + i);
                f0 += b0.x; f1 += b0.y;
              }
              if (col + 63 < N) {
                const float2 b1 = __ldg(reinterpret_cast<const float2*>(ep.bias + col + 32) + i);
                g0 += b1.x; g1 += b1.y;
              }
            }
            if constexpr (EPI == EPI_LNFOLD_BF16) {   // out-of-range columns hold zeros and are clipped by the TMA store
              const uint32_t cb = cs_u32 + (ch * 64 + 2 * i) * 4;
              float c0x, c0y, s0x, s0y, c1x, c1y, s1x, s1y;
              asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(c0x), "=f"(c0y) : "r"(cb));
              asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(s0x), "=f"(s0y) : "r"(cb + BN * 4));
              asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(c1x), "=f"(c1y) : "r"(cb + 128));
              asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(s1x), "=f"(s1y) : "r"(cb + BN * 4 + 128));
              f0 = fmaf(ln_a, f0, fmaf(ln_b, s0x, c0x));
              f1 = fmaf(ln_a, f1, fmaf(ln_b, s0y, c0y));
              g0 = fmaf(ln_a, g0, fmaf(ln_b, s1x, c1x));
              g1 = fmaf(ln_a, g1, fmaf(ln_b, s1y, c1y));
            }
            o[i] = pack_bf16x2(f0, f1);
            o[16 + i] = pack_bf16x2(g0, g1);
          }
          if constexpr (EPI == EPI_BIAS_BF16) {
            if (ep.gn_part != nullptr && col + 63 < N) {   // GroupNorm partials of the rounded values: 16 quads x (sum, sum of squares)
              float v[32];
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                const uint32_t w0 = o[(q < 8 ? 0 : 16) + 2 * (q & 7)], w1 = o[(q < 8 ? 0 : 16) + 2 * (q & 7) + 1];
                const float a = __uint_as_float(w0 << 16), b = __uint_as_float(w0 & 0xffff0000u);
                const float c = __uint_as_float(w1 << 16), d = __uint_as_float(w1 & 0xffff0000u);
                const bool live = row < M;
                v[2 * q] = live ? (a + b) + (c + d) : 0.f;
                v[2 * q + 1] = live ? (a * a + b * b) + (c * c + d * d) : 0.f;
              }
              const float tot = warp_transpose_reduce32(v, lane);   // lane L: quad L / 2, L & 1 ? sum of squares : sum
              ep.gn_part[((size_t)(slab_row >> 5) * (N >> 2) + (col >> 2)) * 2 + lane] = tot;
            }
          }
          uint8_t* slab = slab_acquire();
          stage_row(slab, lane, o);
          slab_release(slab, col, slab_row);
        }
      } else if constexpr (EPI == EPI_BIAS_RESID_LNP) {
        // x_new = x_old + acc + bias written back explicitly (fp32), plus its bf16 copy and the row-statistics partials
#pragma unroll 1
        for (int cp = 0; cp < BN / 64; ++cp) {
          uint32_t ob[32];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int ch = cp * 2 + hh;
            const int col = n0 + ch * 32;
            lnp_advance(tile, ch);
            uint32_t r[32];
            tmem_ld_x32(taddr + ch * 32, r);
            tmem_ld_wait();
            const uint32_t xin = lnp_wait(slab_row, col);
            float ps = 0.f, pq = 0.f;
            const bool live = col + 31 < N;
            const bool have_x = slab_row < M && col < N;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = live ? __ldg(reinterpret_cast<const float4*>(ep.bias + col) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
              const float4 xo = have_x ? lds_v4(xin + ((i ^ (lane & 7)) << 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
              const float w0 = __uint_as_float(r[4 * i]) + b.x + xo.x, w1 = __uint_as_float(r[4 * i + 1]) + b.y + xo.y;
              const float w2 = __uint_as_float(r[4 * i + 2]) + b.z + xo.z, w3 = __uint_as_float(r[4 * i + 3]) + b.w + xo.w;
              ps += (w0 + w1) + (w2 + w3);
              pq += (w0 * w0 + w1 * w1) + (w2 * w2 + w3 * w3);
              ob[hh * 16 + 2 * i] = pack_bf16x2(w0, w1);
              ob[hh * 16 + 2 * i + 1] = pack_bf16x2(w2, w3);
              r[4 * i] = __float_as_uint(w0);
              r[4 * i + 1] = __float_as_uint(w1);
              r[4 * i + 2] = __float_as_uint(w2);
              r[4 * i + 3] = __float_as_uint(w3);
            }
            if (row < M && live) ep.part_out[(size_t)row * (N / 32) + col / 32] = make_float2(ps, pq);
#pragma unroll
            for (int j = 0; j < 8; ++j) sts_v4(xin + ((j ^ (lane & 7)) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
            lnp_store(xin, slab_row, col);
          }
          lnp_store_bf16(ob, slab_row, n0 + cp * 64);
        }
      } else {
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          uint32_t r[32];
          tmem_ld_x32(taddr + ch * 32, r);
          tmem_ld_wait();
          const int col = n0 + ch * 32;
          if constexpr (EPI == EPI_BIAS_RESID_F32) {
            if (col + 31 < N) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col) + i);
                r[4 * i] = __float_as_uint(__uint_as_float(r[4 * i]) + b.x);
                r[4 * i + 1] = __float_as_uint(__uint_as_float(r[4 * i + 1]) + b.y);
                r[4 * i + 2] = __float_as_uint(__uint_as_float(r[4 * i + 2]) + b.z);
                r[4 * i + 3] = __float_as_uint(__uint_as_float(r[4 * i + 3]) + b.w);
              }
            }
          }
          uint8_t* slab = slab_acquire();
          stage_row(slab, lane, r);
          slab_release(slab, col, slab_row);
        }
      }
      // accumulator stage drained: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTAS == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
        else mbar_arrive(&tempty_bar[acc]);
      }
    }
    if (lane == 0) bulk_wait_read<0>();  // staging slabs fully read by the TMA unit before the CTA may exit
  }

  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CTAS == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace tld
