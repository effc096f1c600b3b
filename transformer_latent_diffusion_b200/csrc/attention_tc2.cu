// tcgen05 flash-attention, persistent + pipelined version (head_dim 64, non-causal, fused residual add):
//     x[t, h*64:(h+1)*64] += softmax(q k^T / 8) v            (transformer_blocks.py:31-48,57-59,136)
//
// Why a second kernel: attention_tc.cu runs one (128-query) tile per CTA with single-buffered K/V; its ~8 us per-CTA
// critical path (launch, TMEM alloc, first TMA round trip, serial S -> softmax -> PV chain) is exposed 10x per SM at
// 256 tokens and once per 128-key chunk at 4096 tokens.  Here CTAs are persistent (2 per SM, static round-robin over
// (sample, head, q-tile)), K/V chunks of 64 keys flow through a 2-slot TMA ring, and the roles are decoupled:
//   warp 8  lane 0  TMA producer: Q tile (when the previous tile's last S-MMA has retired) and K|V chunk slots
//   warp 9  lane 0  MMA issuer:   S(c+1) = Q K^T is issued as soon as the softmax threads have pulled S(c) out of TMEM
//                                 (s_free), i.e. it overlaps softmax(c); O(c) = P V after p_full(c); commits release slots
//   warps 0..7      softmax/epilogue, two threads per query row (32 keys each per chunk, S read from TMEM ONCE per chunk):
//                   exchange of the row max through smem (one 256-thread named barrier per chunk), P -> bf16 -> swizzled
//                   K-major smem, O(c) added into registers with the online-softmax correction, finally O/l -> swizzled
//                   fp32 staging -> TMA reduce-add into the residual stream (overlaps the next tile).
// TMEM: S 64 columns + two O accumulators of 64 columns (one per key half).  smem: Q 16 KB + 2 x (K 8 KB + V 8 KB) + P 16 KB + staging 32 KB = 96 KB.
#include "common.h"
#include "launch.h"
#include "ptx.cuh"

#ifdef TLD_TRACE
// developer instrumentation (never compiled into the shipped library): per-role clock64 stamps of CTA 0
__device__ unsigned long long g_t2_trace[3][2048];
#define T2_TR(role, id)                                                                         \
  do {                                                                                          \
    if (blockIdx.x == 0 && trn < 2048) g_t2_trace[role][trn++] = (clock64() << 8) | (id);       \
  } while (0)
extern "C" __attribute__((visibility("default"))) int tld_debug_attention_trace(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_t2_trace, sizeof(g_t2_trace));
}
#else
#define T2_TR(role, id) \
  do {                  \
  } while (0)
#endif

namespace tld {

constexpr int T2_BQ = 128, T2_BK = 64, T2_HD = 64, T2_THREADS = 320;
constexpr int T2_Q_BYTES = T2_BQ * T2_HD * 2;     // 16 KB
constexpr int T2_KV_BYTES = T2_BK * T2_HD * 2;    // 8 KB each for K and V
constexpr int T2_STG_BYTES = 8 * 32 * 128;        // 32 KB: one [32 x 128 B] slab per softmax warp
constexpr int T2_SMEM = 1024 + 2 * T2_Q_BYTES + 2 * 2 * T2_KV_BYTES + T2_STG_BYTES + 256 + 4096;

template <int EMU>
__global__ void __launch_bounds__(T2_THREADS, 2)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     const __grid_constant__ CUtensorMap tmap_x, int n_tok, int D, int B, uint32_t magic_qt,
                     uint32_t magic_h) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2] double-buffered across tiles
  uint8_t* sKV = sQ + 2 * T2_Q_BYTES;                   // slot s: K at sKV + s*16K, V at +8K
  uint8_t* sStg = sKV + 2 * 2 * T2_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + T2_STG_BYTES);
  uint64_t* q_full = bars + 0;    // [2]
  uint64_t* q_empty = bars + 14;  // [2]
  uint64_t* k_full = bars + 2;    // [2]  K and V slots are released separately: K(c) right after S(c) retires, V(c) after
  uint64_t* k_empty = bars + 4;   // [2]  PV(c) - so the K load runs a full softmax period ahead of its S-MMA
  uint64_t* v_full = bars + 6;    // [2]
  uint64_t* v_empty = bars + 8;   // [2]
  uint64_t* s_full = bars + 10;
  uint64_t* s_free = bars + 11;   // 256 arrivals: S(c) is in registers
  uint64_t* p_full = bars + 12;   // [2] 256 arrivals: P(c) written into P buffer c & 1.  One barrier per buffer: the softmax
                                  // threads may run a whole chunk ahead of the MMA thread (they do not depend on V), and a
                                  // single barrier could then complete two phases between two polls of the MMA thread
  uint64_t* o_full = bars + 16;   // [2] PV(c) retired, one barrier per chunk parity: the softmax threads do not wait on every
                                  // phase, and a lone barrier would alias when a thread is two chunks behind
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  float* s_xchg_all = reinterpret_cast<float*>(bars + 32);  // [2 tile parities][m|l][2 halves][128]: alternating buffers
                                                            // make tile k+2's write safe behind tile k+1's barrier

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = D / T2_HD, q_tiles = n_tok / T2_BQ, n_chunks = n_tok / T2_BK;
  const int num_tiles = B * H * q_tiles;

  pdl_launch_dependents();
  if (warp == 8) {
    if (elect_one()) {
      tma_prefetch_desc(&tmap_q);
      tma_prefetch_desc(&tmap_kv);
      tma_prefetch_desc(&tmap_x);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&q_full[s], 1);
        mbar_init(&q_empty[s], 1);
        mbar_init(&k_full[s], 1);
        mbar_init(&k_empty[s], 1);
        mbar_init(&v_full[s], 1);
        mbar_init(&v_empty[s], 1);
      }
      mbar_init(s_full, 1);
      mbar_init(s_free, 256);
      mbar_init(&p_full[0], 256);
      mbar_init(&p_full[1], 256);
      mbar_init(&o_full[0], 1);
      mbar_init(&o_full[1], 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;        // S: columns [0,64) fp32
  const uint32_t tmem_p = tmem_base + 64;   // P: two buffers of 32 columns (64 keys as packed bf16 pairs), [64,96) and [96,128)
  const uint32_t tmem_o = tmem_base + 128;  // O_0 columns [128,192), O_1 columns [192,256)

  // tile -> (sample b, head, q-tile): divisions by q_tiles and H as multiply-high with host-computed ceil(2^32 / d)
  // (exact while tile * d < 2^32; d == 1 is flagged by magic 0) - keeps ~300 cycles of integer division off the tile boundary
  auto tile_coords = [&](int tile, int& row_q, int& row_k, int& head) {
    const uint32_t bh = magic_qt ? __umulhi((uint32_t)tile, magic_qt) : (uint32_t)tile;
    const uint32_t qt = (uint32_t)tile - bh * (uint32_t)q_tiles;
    const uint32_t b = magic_h ? __umulhi(bh, magic_h) : bh;
    head = int(bh - b * (uint32_t)H);
    row_k = int(b) * n_tok;
    row_q = row_k + int(qt) * T2_BQ;
  };

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int it = 0;       // tile counter of this CTA
      int ck = 0;       // global chunk counter (ring position)
      [[maybe_unused]] int trn = 0;
      auto load_q = [&](int tile, int n) {   // n = tile counter of this CTA; Q slot n & 1
        int row_q, row_k, head;
        tile_coords(tile, row_q, row_k, head);
        const int slot = n & 1;
        mbar_wait(&q_empty[slot], ((n >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[slot], T2_Q_BYTES);
        T2_TR(0, 0);
        tma_load_2d(sQ + slot * T2_Q_BYTES, &tmap_q, &q_full[slot], head * T2_HD, row_q);
      };
      load_q(blockIdx.x, 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        int row_q, row_k, head;
        tile_coords(tile, row_q, row_k, head);
        auto load_k = [&](int c, int g) {   // g = global chunk counter
          const int slot = g & 1;
          mbar_wait(&k_empty[slot], ((g >> 1) & 1) ^ 1);
          mbar_expect_tx(&k_full[slot], T2_KV_BYTES);
          T2_TR(0, 1);
          tma_load_2d(sKV + slot * 2 * T2_KV_BYTES, &tmap_kv, &k_full[slot], D + head * T2_HD, row_k + c * T2_BK);
        };
        load_k(0, ck);
        if (tile + (int)gridDim.x < num_tiles) {
          // next tile of this CTA: its Q goes into the other Q buffer a whole tile ahead, and its first K/V boxes are pulled
          // into L2 so that their TMA loads at the tile boundary take an L2 hit instead of a DRAM round trip
          load_q(tile + gridDim.x, it + 1);
          int nq, nk, nh;
          tile_coords(tile + gridDim.x, nq, nk, nh);
          const int npf = n_chunks < 4 ? n_chunks : 4;
          for (int c = 0; c < npf; ++c) {
            tma_prefetch_2d(&tmap_kv, D + nh * T2_HD, nk + c * T2_BK);
            tma_prefetch_2d(&tmap_kv, 2 * D + nh * T2_HD, nk + c * T2_BK);
          }
        }
        for (int c = 0; c < n_chunks; ++c, ++ck) {   // K runs one chunk ahead of V
          if (c + 1 < n_chunks) load_k(c + 1, ck + 1);
          const int slot = ck & 1;
          mbar_wait(&v_empty[slot], ((ck >> 1) & 1) ^ 1);
          mbar_expect_tx(&v_full[slot], T2_KV_BYTES);
          T2_TR(0, 2);
          tma_load_2d(sKV + slot * 2 * T2_KV_BYTES + T2_KV_BYTES, &tmap_kv, &v_full[slot], 2 * D + head * T2_HD,
                      row_k + c * T2_BK);
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(T2_BQ, T2_BK, 0, 0);   // S = Q K^T : both K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(T2_BQ, T2_HD, 0, 1);   // O = P V   : V MN-major
      const uint64_t qdesc0 = umma_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      [[maybe_unused]] int trn = 0;
      auto issue_s = [&](int slot, int qslot) {
        T2_TR(1, 1);
        const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(sKV + slot * 2 * T2_KV_BYTES), 16, 1024);
        const uint64_t qdesc = qdesc0 + uint64_t(qslot * (T2_Q_BYTES >> 4));
#pragma unroll
        for (int k = 0; k < T2_HD / 16; ++k) umma_ss_f16(tmem_s, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
        umma_commit(&k_empty[slot]);
      };
      // One flat sequence of chunks g = 0 .. G-1 over all tiles of this CTA: S(g+1) is always issued before the wait for
      // P(g), also across a tile boundary (the next tile's first S overlaps the last softmax of the current tile).
      const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int G = my_tiles * n_chunks;
      int c = 0, it = 0;  // chunk within the tile, tile counter of S(g)
      mbar_wait(&q_full[0], 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      for (int g = 0; g < G; ++g) {
        if (g + 1 < G) {
          const int c1 = (c + 1 == n_chunks) ? 0 : c + 1;
          const int it1 = it + (c1 == 0);   // tile counter of S(g+1)
          if (c1 == 0) mbar_wait(&q_full[it1 & 1], (it1 >> 1) & 1);
          mbar_wait(&k_full[(g + 1) & 1], ((g + 1) >> 1) & 1);
          mbar_wait(s_free, g & 1);   // S(g) is in the softmax threads' registers
          tc_fence_after();
          issue_s((g + 1) & 1, it1 & 1);
          if (c1 == n_chunks - 1) umma_commit(&q_empty[it1 & 1]);   // the tile's last S-MMA is queued
        }
        const int slot = g & 1;
        mbar_wait(&v_full[slot], (g >> 1) & 1);
        mbar_wait(&p_full[slot], (g >> 1) & 1);
        tc_fence_after();
        T2_TR(1, 2);
        const uint32_t vbase = smem_u32(sKV + slot * 2 * T2_KV_BYTES + T2_KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < T2_BK / 16; ++kk) {   // keys [16 kk, +16) belong to half kk >> 1 -> accumulator O_(kk>>1)
          const uint64_t bdesc = umma_smem_desc_sw128(vbase + kk * 16 * 128, T2_BK * 128, 1024);
          // A = P(g) straight from TMEM (buffer g & 1, 8 columns of packed bf16 pairs per K=16 step)
          umma_ts_f16(tmem_o + (kk >> 1) * 64, tmem_p + slot * 32 + kk * 8, bdesc, idesc_o, (c | (kk & 1)) != 0);
        }
        umma_commit(&o_full[slot]);
        umma_commit(&v_empty[slot]);
        if (++c == n_chunks) {
          c = 0;
          ++it;
        }
      }
    }
  } else {
    // ===================== softmax + epilogue (two threads per query row, independent key halves) =====================
    // Thread (hf, r) owns keys [32 hf, 32 hf + 32) of every chunk for query row r: its own running reference m_ref, row
    // sum l and its own accumulator O_hf in TMEM (the P V product of a chunk is issued as two K=32 halves into O_0 / O_1),
    // so the two threads of a row never talk inside the key loop; (m, l) are exchanged ONCE per tile and the halves are
    // merged in the epilogue:  out = (a_0 O_0 + a_1 O_1) / (a_0 l_0 + a_1 l_1),  a_h = 2^(m_h - max(m_0, m_1)).
    // m_ref moves only when the half-row maximum grew by more than 2^8 in the exp2 domain (lazy rescaling: P <= 256 is
    // harmless in bf16/fp32); only then is O_hf pulled out of TMEM, scaled and written back - rare after chunk 0.
    // EMU of the 16 exp2 pairs per thread and chunk run as a polynomial on the FMA pipe (exp2_fma2), the rest on MUFU.
    const int r = threadIdx.x & 127, hf = threadIdx.x >> 7;
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    const uint32_t tmem_mine = tmem_o + lane_base + hf * 64, tmem_other = tmem_o + lane_base + (hf ^ 1) * 64;
    const float sl2 = 0.125f * 1.4426950408889634f;
    const uint32_t xchg0 = smem_u32(s_xchg_all);
    uint32_t gs = 0, it = 0;
    [[maybe_unused]] int trn = threadIdx.x == 0 ? 0 : 4096;
    // tile epilogue (merge the two key halves, x += O / l); runs AFTER the exp2 work of the next tile's first chunk so that
    // the latency of the last P V product is hidden (O_0/O_1 stay intact until that chunk's p_full arrival)
    auto epilogue = [&](float m_fin, float l_fin, int row_q, int head, uint32_t g_last, uint32_t xpar) {
      const uint32_t xb = xchg0 + xpar * 2048;   // [2 tile parities][m|l][2 halves][128 rows]
      sts_f32(xb + (hf * 128 + r) * 4, m_fin);
      sts_f32(xb + 1024 + (hf * 128 + r) * 4, l_fin);
      named_bar_sync(1, 256);
      const float m_p = lds_f32(xb + ((hf ^ 1) * 128 + r) * 4), l_p = lds_f32(xb + 1024 + ((hf ^ 1) * 128 + r) * 4);
      const float m_all = fmaxf(m_fin, m_p);
      float a_me = exp2f((m_fin - m_all) * sl2), a_p = exp2f((m_p - m_all) * sl2);
      const float inv = 1.f / (l_fin * a_me + l_p * a_p);
      a_me *= inv;
      a_p *= inv;
      const uint32_t slab = smem_u32(sStg) + warp * (32 * 128);
      if (lane == 0) bulk_wait_read<0>();  // the previous tile's reduce-add has finished reading this slab
      __syncwarp();
      mbar_wait(&o_full[g_last & 1], (g_last >> 1) & 1);   // the tile's last P V product has retired
      tc_fence_after();
      T2_TR(2, 3);
#pragma unroll
      for (int j = 0; j < 2; ++j) {   // output columns [32 hf + 16 j, +16) from both accumulators
        uint32_t o0[16], o1[16];
        tmem_ld_x16(tmem_mine + hf * 32 + j * 16, o0);
        tmem_ld_x16(tmem_other + hf * 32 + j * 16, o1);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            v[e] = __uint_as_float(o0[4 * q + e]) * a_me + __uint_as_float(o1[4 * q + e]) * a_p;
          sts_v4(slab + lane * 128 + (((j * 4 + q) ^ (lane & 7)) << 4), __float_as_uint(v[0]), __float_as_uint(v[1]),
                 __float_as_uint(v[2]), __float_as_uint(v[3]));
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_2d(&tmap_x, reinterpret_cast<void*>(sStg + warp * (32 * 128)), head * T2_HD + hf * 32,
                          row_q + (warp & 3) * 32);
        bulk_commit();
      }
      T2_TR(2, 4);
    };
    bool pending = false;
    float m_prev = 0.f, l_prev = 0.f;
    int row_q_prev = 0, head_prev = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int row_q, row_k, head;
      tile_coords(tile, row_q, row_k, head);
      T2_TR(2, 0);
      float m_ref = -INFINITY, l_run = 0.f;
      for (int c = 0; c < n_chunks; ++c, ++gs) {
        const uint32_t ph = gs & 1;
        mbar_wait(s_full, ph);
        tc_fence_after();
        T2_TR(2, 1);
        uint32_t s[32];
        tmem_ld_x32(tmem_s + lane_base + hf * 32, s);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_free);  // S(c) is in registers: the tensor core may overwrite it with S(c+1)
        float m0 = __uint_as_float(s[0]), m1 = __uint_as_float(s[1]);
#pragma unroll
        for (int i = 2; i < 30; i += 4) {
          m0 = fmax3(m0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
          m1 = fmax3(m1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        }
        const float mloc = fmax3(fmaxf(m0, m1), __uint_as_float(s[30]), __uint_as_float(s[31]));
        const bool need = (mloc - m_ref) * sl2 > 8.f;
        const bool fix_o = __any_sync(0xffffffffu, need) && c > 0;
        float corr = 1.f;
        if (need) {
          corr = exp2f((m_ref - mloc) * sl2);
          m_ref = mloc;
        }
        l_run *= corr;
        if (fix_o) {  // rare: rescale this thread's O_hf in TMEM (no PV is in flight between o_full(c-1) and p_full(c))
          mbar_wait(&o_full[(gs - 1) & 1], ((gs - 1) >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t v[16];
            tmem_ld_x16(tmem_mine + j * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * corr);
            tmem_st_x16(tmem_mine + j * 16, v);
          }
          tmem_st_wait();
          tc_fence_before();
        }
        const float mb = m_ref * sl2;
        uint32_t pk[16];
        float2 rs2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 a = ffma2(make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])),
                                 make_float2(sl2, sl2), make_float2(-mb, -mb));
          const bool emulate = (i * EMU) / 16 != ((i + 1) * EMU) / 16;   // EMU of 16 pairs, evenly interleaved
          const float2 p = emulate ? exp2_fma2(a) : make_float2(ex2_approx(a.x), ex2_approx(a.y));
          rs2 = fadd2(rs2, p);
          pk[i] = pack_bf16x2(p.x, p.y);
        }
        l_run += rs2.x + rs2.y;
        // S(c) retired => every earlier MMA retired, in particular PV(c-2), the last reader of P buffer c & 1
        if (c == 0 && pending) epilogue(m_prev, l_prev, row_q_prev, head_prev, gs - 1, (it & 1) ^ 1);
        tmem_st_x16(tmem_p + lane_base + (gs & 1) * 32 + hf * 16, pk);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[gs & 1]);
        T2_TR(2, 2);
      }
      pending = true;
      m_prev = m_ref;
      l_prev = l_run;
      row_q_prev = row_q;
      head_prev = head;
    }
    if (pending) epilogue(m_prev, l_prev, row_q_prev, head_prev, gs - 1, (it & 1) ^ 1);
    if (lane == 0) bulk_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static int g_exp_emu = 6;
void set_attention_exp_emu(int v) { g_exp_emu = v; }

template <int EMU>
static int launch_tc2(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& tx, long long tiles, int n_tok, int D,
                      int B, cudaStream_t st) {
  auto magic = [](uint32_t d) -> uint32_t { return d == 1 ? 0u : uint32_t(((1ull << 32) + d - 1) / d); };
  const uint32_t magic_qt = magic(n_tok / T2_BQ), magic_h = magic(D / T2_HD);
  static bool attr_set = false;
  if (!attr_set) {
    TLD_CUDA_OK(cudaFuncSetAttribute(attention_tc2_kernel<EMU>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM));
    attr_set = true;
  }
  long long grid = 2LL * sm_count();
  if (grid > tiles) grid = tiles;
  return launch_pdl(attention_tc2_kernel<EMU>, dim3((unsigned)grid), dim3(T2_THREADS), T2_SMEM, st, tq, tkv, tx, n_tok, D, B,
                    magic_qt, magic_h);
}

int launch_self_attention_tc2(const bf16* qkv, float* x, int B, int n_tok, int D, cudaStream_t st) {
  TLD_CHECK(D % 64 == 0 && n_tok % 128 == 0, "attention_tc2: needs embed_dim % 64 == 0 and tokens per sample % 128 == 0");
  const long long T = (long long)B * n_tok;
  CUtensorMap tq, tkv, tx;
  if (make_tmap_2d(&tq, qkv, false, T, 3LL * D, 3LL * D, T2_BQ)) return 1;   // box 128 rows x 64 cols
  if (make_tmap_2d(&tkv, qkv, false, T, 3LL * D, 3LL * D, T2_BK)) return 1;  // box  64 rows x 64 cols
  if (make_tmap_2d(&tx, x, true, T, D, D, 32)) return 1;
  const long long tiles = (long long)B * (D / 64) * (n_tok / T2_BQ);
  TLD_CHECK(tiles * (D / 64 > n_tok / T2_BQ ? D / 64 : n_tok / T2_BQ) < (1ll << 32),
            "attention_tc2: too many tiles for the 32-bit multiply-high tile decomposition");
  switch (g_exp_emu) {
    case 0: return launch_tc2<0>(tq, tkv, tx, tiles, n_tok, D, B, st);
    case 4: return launch_tc2<4>(tq, tkv, tx, tiles, n_tok, D, B, st);
    case 6: return launch_tc2<6>(tq, tkv, tx, tiles, n_tok, D, B, st);
    case 8: return launch_tc2<8>(tq, tkv, tx, tiles, n_tok, D, B, st);
    case 10: return launch_tc2<10>(tq, tkv, tx, tiles, n_tok, D, B, st);
    default: TLD_CHECK(false, "attention exp emulation: pairs per 16 must be one of 0, 4, 6, 8, 10");
  }
}

}  // namespace tld
