// SelfAttention of one decoder block as ONE kernel for the 16 x 16 token grid (reference tld/transformer_blocks.py:51-59 =
// qkv_linear + MHAttention, :24-48 = scaled_dot_product_attention, :136 = the residual add):
//     x[t, h*64:(h+1)*64] += softmax(q_h k_h^T / 8) v_h,   [q | k | v] = LN1(x) W_qkv^T          (no bias, no out-projection)
//
// The separate kernels (qkv GEMM, attention_tc2) write the 3D-wide qkv tensor (151 MB per layer at T = 32768) and read it
// back (251 MB: K and V once per 128-query tile).  Here a CTA pair owns one (sample, head) at a time: 256 tokens = the pair's
// M = 256 rows, so q, k, v of the head never leave the two SMs.
//
//   G    acc[256 x 192] = xn[256 x D] . W_h[192 x D]^T      tcgen05.mma.cta_group::2, M 256, N 192 (q | k | v of head h),
//                                                           TMA -> 128B-swizzled smem ring, exactly the pair GEMM's main loop
//   drain  acc (TMEM) -> bf16 -> this CTA's shared memory:  Q [128 own tokens][64] and K [128 own keys][64] K-major, and V
//        TRANSPOSED: the pair MMA splits B's N over the two CTAs, so for O = P V (N = head_dim) CTA c must hold d-half c of V
//        for ALL 256 keys: Vt [32 d][256 keys] K-major.  A thread owns one key; three rounds of warp shuffles (xor 1, 2, 4)
//        turn 32 d-values of 8 adjacent keys into 16-byte units (one d, 8 keys); the d-half of the peer goes straight into
//        the peer's Vt through distributed shared memory (st.async ... mbarrier::complete_tx::bytes on the peer's barrier).
//   S    S[256 x 256] = Q K^T                               M 256, N 256 (each CTA contributes its 128 keys), K 64
//   softmax  two threads per query row (128 keys each), S read from TMEM once into registers, row max exchanged through
//        shared memory, P = 2^((s - m) / 8 log2 e) as packed bf16 written back over the S columns (P aliases S)
//   PV   O[256 x 64] = P V                                  A = P from TMEM, B = Vt, M 256, N 64, K 256
//   epilogue  O / l -> swizzled fp32 slab -> TMA reduce-add into the residual stream (as attention_tc2)
//
// TMEM (512 columns, all of it): acc [0,192), O [192,256), S/P [256,512).  The accumulator is single-buffered, so the
// tensor pipe order per item i is  G(i) . PV(i-1) . S(i) . G(i+1) ...: softmax(i) runs under G(i+1), and PV(i-1) is issued
// behind G(i) (its P has long been ready) so that it executes while the drain threads pull Q(i) / K(i) out of the
// accumulator - S(i) and G(i+1) follow without a bubble.  Vt(i) is written only after PV(i-1) has retired (o_full).
//
// Warps: 0 TMA producer, 1 MMA issuer (leader CTA), 2 TMEM allocator, 3 forwarder (tells the leader's MMA thread that THIS
// CTA's Q/K/Vt - including the peer's pushed half - are complete), 4..11 drain + softmax + epilogue.
// Roofline: tensor (2 . 256 . 192 . D + 2 . 2 . 256 . 256 . 64 flop per item).  Algorithmic bytes per layer at T = 32768, D = 768:
// xn 50 MB + x read-modify-write 200 MB (the separate kernels: + 151 MB qkv written + 251 MB read).
#include <cudaTypedefs.h>

#include "common.h"
#include "launch.h"
#include "ptx.cuh"

#ifdef TLD_TRACE
// developer instrumentation (never compiled into the shipped library): per-role clock64 stamps of CTA 0
__device__ unsigned long long g_qa_trace[4][2048];
#define QA_TR(role, id)                                                                       \
  do {                                                                                        \
    if (blockIdx.x == 0 && trn < 2048) g_qa_trace[role][trn++] = (clock64() << 8) | (id);     \
  } while (0)
extern "C" __attribute__((visibility("default"))) int tld_debug_qa_trace(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_qa_trace, sizeof(g_qa_trace));
}
#else
#define QA_TR(role, id) \
  do {                  \
  } while (0)
#endif

namespace tld {

constexpr int QA_BK = 64, QA_STAGES = 5, QA_THREADS = 384;
constexpr int QA_A_BYTES = 128 * QA_BK * 2;   // this CTA's 128 tokens
constexpr int QA_B_BYTES = 96 * QA_BK * 2;    // this CTA's half of the 192 weight rows
constexpr int QA_STAGE_BYTES = QA_A_BYTES + QA_B_BYTES;
constexpr int QA_QKV_BYTES = 3 * 16384;       // Q, K, Vt
constexpr int QA_SLAB_BYTES = 8 * 4096;       // one [32 rows x 128 B] fp32 slab per softmax warp
constexpr int QA_XCHG_BYTES = 2048;           // m[2][128], l[2][128]
constexpr int QA_SMEM = 1024 + QA_STAGES * QA_STAGE_BYTES + QA_QKV_BYTES + QA_SLAB_BYTES + QA_XCHG_BYTES + 256;
constexpr uint32_t QA_COL_O = 192, QA_COL_S = 256;

__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, uint32_t cluster_bar, uint32_t a, uint32_t b, uint32_t c,
                                            uint32_t d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(cluster_addr),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(cluster_bar)
               : "memory");
}
// O(+)= P V with P in TMEM (A operand) for the CTA pair
__device__ __forceinline__ void umma_ts_f16_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 32 values (d = 0..31) of this lane's key -> four 16-byte units: unit j3 = d' = 8 j3 + (lane & 7), keys of the lane's
// 8-lane group in ascending order (bf16).  Rounds: xor 1 pairs keys, xor 2 makes 4-key words, xor 4 makes 8-key units.
__device__ __forceinline__ void transpose_keys8(const uint32_t (&ra)[32], int lane, uint4 (&out)[4]) {
  const bool p0 = lane & 1, p1 = lane & 2, p2 = lane & 4;
  uint32_t w1[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float a = __uint_as_float(ra[2 * j]), b = __uint_as_float(ra[2 * j + 1]);
    const float recv = __shfl_xor_sync(0xffffffffu, p0 ? a : b, 1);
    w1[j] = p0 ? pack_bf16x2(recv, b) : pack_bf16x2(a, recv);
  }
  uint32_t w2[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t keep = p1 ? w1[2 * j + 1] : w1[2 * j];
    const uint32_t recv = __shfl_xor_sync(0xffffffffu, p1 ? w1[2 * j] : w1[2 * j + 1], 2);
    w2[j][0] = p1 ? recv : keep;
    w2[j][1] = p1 ? keep : recv;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t k0 = p2 ? w2[2 * j + 1][0] : w2[2 * j][0], k1 = p2 ? w2[2 * j + 1][1] : w2[2 * j][1];
    const uint32_t r0 = __shfl_xor_sync(0xffffffffu, p2 ? w2[2 * j][0] : w2[2 * j + 1][0], 4);
    const uint32_t r1 = __shfl_xor_sync(0xffffffffu, p2 ? w2[2 * j][1] : w2[2 * j + 1][1], 4);
    out[j] = p2 ? make_uint4(r0, r1, k0, k1) : make_uint4(k0, k1, r0, r1);
  }
}

template <int EMU>
__global__ void __launch_bounds__(QA_THREADS, 1)
qkv_attention_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                     const __grid_constant__ CUtensorMap tmap_x, int n_items, int D) {
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  constexpr int STAGES = QA_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * QA_A_BYTES;
  uint8_t* sQ = smem + STAGES * QA_STAGE_BYTES;
  uint8_t* sK = sQ + 16384;
  uint8_t* sVt = sK + 16384;
  uint8_t* sSlab = sVt + 16384;
  float* xchg = reinterpret_cast<float*>(sSlab + QA_SLAB_BYTES);   // m[2][128] | l[2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSlab + QA_SLAB_BYTES + QA_XCHG_BYTES);
  uint64_t* full_bar = bars;               // [STAGES] (leader's is used: both CTAs' TMA bytes land there)
  uint64_t* empty_bar = bars + STAGES;     // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;      // G(i) retired (multicast commit)
  uint64_t* vt_ready = bars + 2 * STAGES + 1;  // leader's: both forwarders arrived (Vt of both CTAs complete)
  uint64_t* drain_done = bars + 2 * STAGES + 2;  // 8 local drain warps + 8 KB of Vt pushed by the peer
  uint64_t* s_full = bars + 2 * STAGES + 3;    // S(i) retired (multicast commit)
  uint64_t* p_full = bars + 2 * STAGES + 4;    // leader's: 16 softmax warps of the pair wrote P(i)
  uint64_t* qk_ready = bars + 2 * STAGES + 5;  // leader's: 16 drain warps stored Q / K and hold their v chunk (accumulator free)
  uint64_t* o_full = bars + 2 * STAGES + 6;    // PV(i) retired (multicast commit): O(i) complete, Vt and P free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = D / 64, k_blocks = D / QA_BK;
  const int first_item = blockIdx.x / 2, item_step = gridDim.x / 2;

  pdl_launch_dependents();
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(vt_ready, 2);
    mbar_init(qk_ready, 16);
    mbar_init(o_full, 1);
    mbar_init(drain_done, 8);
    mbar_init(s_full, 1);
    mbar_init(p_full, 16);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
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
      for (int item = first_item; item < n_items; item += item_step) {
        const int b = item / H, h = item - b * H;
        const int m0 = b * 256 + int(cta_rank) * 128;
        int wrow[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {   // rows [96 rank, +96) of the head's 192 (q | k | v) weight rows, 32 at a time
          const int n = int(cta_rank) * 96 + 32 * j;
          wrow[j] = (n >> 6) * D + h * 64 + (n & 63);
        }
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full_bar[stage], QA_STAGE_BYTES * 2);
          tma_load_2d_pair(smem_a + stage * QA_A_BYTES, &tmap_a, &full_bar[stage], kb * QA_BK, m0);
#pragma unroll
          for (int j = 0; j < 3; ++j)
            tma_load_2d_pair(smem_b + stage * QA_B_BYTES + j * 4096, &tmap_w, &full_bar[stage], kb * QA_BK, wrow[j]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (leader && elect_one()) {
      constexpr uint32_t idesc_g = umma_idesc_bf16(256, 192, 0, 0);
      constexpr uint32_t idesc_s = umma_idesc_bf16(256, 256, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(256, 64, 0, 0);
      const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(sQ), 16u, 1024);
      const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(sK), 16u, 1024);
      const uint32_t vt_u32 = smem_u32(sVt);
      auto issue_pv = [&]() {
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {   // 16 keys per step: 8 packed P columns, Vt atom kk / 4 (64 keys), +32 B per step
          const uint64_t bdesc = umma_smem_desc_sw128(vt_u32 + (kk >> 2) * 4096 + (kk & 3) * 32, 16u, 1024);
          umma_ts_f16_pair(tmem_base + QA_COL_O, tmem_base + QA_COL_S + kk * 8, bdesc, idesc_o, kk != 0);
        }
      };
      int stage = 0;
      uint32_t phase = 0, it = 0;
      [[maybe_unused]] int trn = 0;
      for (int item = first_item; item < n_items; item += item_step, ++it) {
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (kb == 0) QA_TR(0, 1);
          tc_fence_after();
          const uint64_t adesc = umma_smem_desc_sw128(smem_u32(smem_a + stage * QA_A_BYTES), 16u, 1024);
          const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(smem_b + stage * QA_B_BYTES), 16u, 1024);
#pragma unroll
          for (int k = 0; k < QA_BK / 16; ++k) umma_ss_f16_pair(tmem_base, adesc + 2u * k, bdesc + 2u * k, idesc_g, (kb | k) != 0);
          umma_commit_pair(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        QA_TR(0, 2);
        umma_commit_pair(acc_full);
        if (it > 0) {   // PV(i-1) behind G(i): it executes while the drain of acc(i) runs, so the tensor pipe does not idle
          mbar_wait(vt_ready, (it - 1) & 1);
          mbar_wait(p_full, (it - 1) & 1);
          QA_TR(0, 3);
          tc_fence_after();
          issue_pv();
          umma_commit_pair(o_full);
        }
        mbar_wait(qk_ready, it & 1);
        QA_TR(0, 4);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss_f16_pair(tmem_base + QA_COL_S, qdesc + 2u * k, kdesc + 2u * k, idesc_s, k != 0);
        umma_commit_pair(s_full);
      }
      if (it > 0) {
        mbar_wait(vt_ready, (it - 1) & 1);
        mbar_wait(p_full, (it - 1) & 1);
        tc_fence_after();
        issue_pv();
        umma_commit_pair(o_full);
      }
    }
  } else if (warp == 3) {
    // ===================== forwarder: this CTA's Q / K / Vt are complete -> leader's qkv_ready =====================
    if (elect_one()) {
      const uint32_t remote = mapa_u32(smem_u32(vt_ready), 0);
      uint32_t it = 0;
      [[maybe_unused]] int trn = 0;
      for (int item = first_item; item < n_items; item += item_step, ++it) {
        mbar_wait(drain_done, it & 1);
        QA_TR(2, 1);
        fence_proxy_async_smem();   // the peer's st.async bytes in our Vt -> visible to the tensor core's (async proxy) reads
        QA_TR(2, 2);
        mbar_arrive_cluster(remote);
      }
    }
  } else if (warp >= 4) {
    // ===================== drain + softmax + epilogue =====================
    const int sw = warp - 4, qd = sw & 3, hf = sw >> 2;
    const int r = qd * 32 + lane;                      // token of this CTA (TMEM lane)
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    const uint32_t t_acc = tmem_base + lane_base, t_o = tmem_base + lane_base + QA_COL_O, t_s = tmem_base + lane_base + QA_COL_S;
    const uint32_t q_u32 = smem_u32(sQ), k_u32 = smem_u32(sK), vt_u32 = smem_u32(sVt);
    const uint32_t xm = smem_u32(xchg), xl = xm + 1024;
    const uint32_t slab = smem_u32(sSlab) + sw * 4096;
    const uint32_t peer = cta_rank ^ 1u;
    const uint32_t peer_bar = mapa_u32(smem_u32(drain_done), peer);
    const uint32_t qk_remote = mapa_u32(smem_u32(qk_ready), 0);
    const float sl2 = 0.125f * 1.4426950408889634f;
    // Vt unit address of this lane: keys [key0, key0 + 8) of the pair's 256 -> atom key0 / 64, 16-byte chunk (key0 % 64) / 8
    const int key0 = int(cta_rank) * 128 + qd * 32 + (lane & ~7);
    const uint32_t vt_unit = vt_u32 + (key0 >> 6) * 4096 + (lane & 7) * 128 + ((((key0 & 63) >> 3) ^ (lane & 7)) << 4);

    auto store_rows = [&](uint32_t dst_row, int half, const uint32_t (&ra)[32]) {   // 32 fp32 -> bf16 -> 4 chunks of a 128 B row
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sts_v4(dst_row + (((half * 4 + j) ^ (lane & 7)) << 4),
               pack_bf16x2(__uint_as_float(ra[8 * j]), __uint_as_float(ra[8 * j + 1])),
               pack_bf16x2(__uint_as_float(ra[8 * j + 2]), __uint_as_float(ra[8 * j + 3])),
               pack_bf16x2(__uint_as_float(ra[8 * j + 4]), __uint_as_float(ra[8 * j + 5])),
               pack_bf16x2(__uint_as_float(ra[8 * j + 6]), __uint_as_float(ra[8 * j + 7])));
    };
    auto epilogue = [&](int row0, int head) {   // x[row0 + r, head * 64 + hf * 32 .. +32) += O / l
      const float inv = 1.f / (lds_f32(xl + r * 4) + lds_f32(xl + 512 + r * 4));
      if (lane == 0) bulk_wait_read<0>();   // the previous reduce-add has finished reading this slab
      __syncwarp();
      uint32_t o[32];
      tmem_ld_x32(t_o + hf * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts_v4(slab + lane * 128 + ((j ^ (lane & 7)) << 4), __float_as_uint(__uint_as_float(o[4 * j]) * inv),
               __float_as_uint(__uint_as_float(o[4 * j + 1]) * inv), __float_as_uint(__uint_as_float(o[4 * j + 2]) * inv),
               __float_as_uint(__uint_as_float(o[4 * j + 3]) * inv));
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_2d(&tmap_x, reinterpret_cast<void*>(sSlab + sw * 4096), head * 64 + hf * 32, row0 + qd * 32);
        bulk_commit();
      }
    };

    uint32_t it = 0;
    int row0_prev = 0, head_prev = 0;
    [[maybe_unused]] int trn = (threadIdx.x == 128 || threadIdx.x == 256) ? 0 : 4096;
    [[maybe_unused]] const int trole = threadIdx.x == 128 ? 1 : 3;
    for (int item = first_item; item < n_items; item += item_step, ++it) {
      const int b = item / H, h = item - b * H;
      const int row0 = b * 256 + int(cta_rank) * 128;
      mbar_wait(acc_full, it & 1);
      QA_TR(trole, 1);
      tc_fence_after();
      // ---- drain: thread hf = 0 takes q and d-half 0 of v, hf = 1 takes k and d-half 1 of v.  Q and K go first and the v
      // chunk is only pulled into registers: then the accumulator is free and S(i) / G(i+1) can be issued while the
      // transposes (off the tensor pipe's critical path: Vt is first needed by PV(i), a whole G later) run.
      {
        uint32_t ra[32], rb[32], rv[32];
        const uint32_t dst = (hf == 0 ? q_u32 : k_u32) + r * 128;
        tmem_ld_x32(t_acc + hf * 64, ra);
        tmem_ld_x32(t_acc + hf * 64 + 32, rb);
        tmem_ld_x32(t_acc + 128 + hf * 32, rv);
        tmem_ld_wait();
        store_rows(dst, 0, ra);
        store_rows(dst, 1, rb);
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(qk_remote);
        QA_TR(trole, 6);
        if (it > 0) {   // PV(i-1) retired: O(i-1) complete, and Vt (both CTAs': the commit is multicast) may be overwritten
          mbar_wait(o_full, (it - 1) & 1);
          QA_TR(trole, 7);
          tc_fence_after();
        }
        uint4 u[4];
        transpose_keys8(rv, lane, u);
        if (uint32_t(hf) == cta_rank) {   // d-half hf of v belongs to CTA hf's Vt
#pragma unroll
          for (int j = 0; j < 4; ++j) sts_v4(vt_unit + j * 1024, u[j].x, u[j].y, u[j].z, u[j].w);
        } else {
          const uint32_t remote = mapa_u32(vt_unit, peer);
#pragma unroll
          for (int j = 0; j < 4; ++j) st_async_v4(remote + j * 1024, peer_bar, u[j].x, u[j].y, u[j].z, u[j].w);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (sw == 0) mbar_expect_tx(drain_done, 8192);   // arrive + the 8 KB of Vt the peer pushes in this phase
        else mbar_arrive(drain_done);
      }
      QA_TR(trole, 2);
      if (it > 0) epilogue(row0_prev, head_prev);   // O(i-1): PV(i-1) retired together with G(i)
      row0_prev = row0;
      head_prev = h;
      // ---- softmax of row r over keys [128 hf, +128)
      QA_TR(trole, 3);
      mbar_wait(s_full, it & 1);
      QA_TR(trole, 4);
      tc_fence_after();
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(t_s + hf * 128 + c * 32, reinterpret_cast<uint32_t(&)[32]>(s[c * 32]));
      tmem_ld_wait();
      float m0 = __uint_as_float(s[0]), m1 = __uint_as_float(s[1]);
#pragma unroll
      for (int i = 2; i < 126; i += 4) {
        m0 = fmax3(m0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
        m1 = fmax3(m1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
      }
      const float mloc = fmax3(fmaxf(m0, m1), __uint_as_float(s[126]), __uint_as_float(s[127]));
      sts_f32(xm + (hf * 128 + r) * 4, mloc);
      tc_fence_before();
      named_bar_sync(1, 256);   // both threads of every row hold their S in registers: P may overwrite S
      tc_fence_after();
      const float mb = fmaxf(mloc, lds_f32(xm + ((hf ^ 1) * 128 + r) * 4)) * sl2;
      float2 rs2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 a = ffma2(make_float2(__uint_as_float(s[c * 32 + 2 * i]), __uint_as_float(s[c * 32 + 2 * i + 1])),
                                 make_float2(sl2, sl2), make_float2(-mb, -mb));
          const bool emulate = (i * EMU) / 16 != ((i + 1) * EMU) / 16;
          const float2 p = emulate ? exp2_fma2(a) : make_float2(ex2_approx(a.x), ex2_approx(a.y));
          rs2 = fadd2(rs2, p);
          pk[i] = pack_bf16x2(p.x, p.y);
        }
        tmem_st_x16(t_s + hf * 64 + c * 16, pk);
      }
      tmem_st_wait();
      sts_f32(xl + (hf * 128 + r) * 4, rs2.x + rs2.y);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(p_full), 0));
      QA_TR(trole, 5);
    }
    if (it > 0) {
      mbar_wait(o_full, (it - 1) & 1);
      tc_fence_after();
      epilogue(row0_prev, head_prev);
    }
    if (lane == 0) bulk_wait_read<0>();
  }

  tc_fence_before();
  cluster_sync_all();   // nobody exits while the peer may still push into our Vt or arrive on one of our barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

static int g_qa_emu = 6;
void set_qkv_attention_exp_emu(int v) { g_qa_emu = v; }

template <int EMU>
static int launch_qa(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tx, int n_items, int D, cudaStream_t st) {
  auto kern = qkv_attention_kernel<EMU>;
  static bool attr_set = false;
  if (!attr_set) {
    TLD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, QA_SMEM));
    attr_set = true;
  }
  const int slots = sm_count() / 2;
  const int grid = (n_items < slots ? n_items : slots) * 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(QA_THREADS);
  cfg.dynamicSmemBytes = QA_SMEM;
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
  TLD_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tw, tx, n_items, D));
  return 0;
}

// x[B * 256, D] (fp32) += SelfAttention(xn[B * 256, D] (bf16), Wqkv[3 D, D] (bf16)) for 256 tokens per sample, head_dim 64
int launch_qkv_attention(const bf16* xn, const bf16* wqkv, float* x, int B, int n_tok, int D, cudaStream_t st) {
  TLD_CHECK(B > 0 && n_tok == 256, "qkv_attention: the fused kernel covers exactly 256 tokens per sample (16 x 16 grid)");
  TLD_CHECK(D > 0 && D % 64 == 0, "qkv_attention: embed_dim must be a multiple of 64 (head_dim 64)");
  TLD_CHECK(((reinterpret_cast<uintptr_t>(xn) | reinterpret_cast<uintptr_t>(wqkv) | reinterpret_cast<uintptr_t>(x)) & 15) == 0,
            "qkv_attention: operands must be 16-byte aligned");
  const long long T = (long long)B * 256;
  CUtensorMap ta, tw, tx;
  if (make_tmap_2d(&ta, xn, false, T, D, D, 128)) return 1;
  if (make_tmap_2d(&tw, wqkv, false, 3LL * D, D, D, 32)) return 1;
  if (make_tmap_2d(&tx, x, true, T, D, D, 32)) return 1;
  const long long items = (long long)B * (D / 64);
  TLD_CHECK(items < (1ll << 31), "qkv_attention: too many (sample, head) items");
  switch (g_qa_emu) {
    case 0: return launch_qa<0>(ta, tw, tx, (int)items, D, st);
    case 4: return launch_qa<4>(ta, tw, tx, (int)items, D, st);
    case 6: return launch_qa<6>(ta, tw, tx, (int)items, D, st);
    case 8: return launch_qa<8>(ta, tw, tx, (int)items, D, st);
    default: TLD_CHECK(false, "qkv_attention exp emulation: pairs per 16 must be one of 0, 4, 6, 8");
  }
}

}  // namespace tld
