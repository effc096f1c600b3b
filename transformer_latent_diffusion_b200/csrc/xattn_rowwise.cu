// norm2 + CrossAttention + residual + norm3 of a decoder block as ONE row-wise kernel (reference tld/transformer_blocks.py:
// 62-75 CrossAttention, :137 x = cross_attention(norm2(x), y) + x, :138 norm3):
//
// The conditioning sequence has exactly TWO tokens (noise level, label: tld/denoiser.py:118-123), so the attention of a query
// row over them is a softmax over two logits per head - a sigmoid of their difference:
//     p0 = sigmoid((q_h . k0_h - q_h . k1_h) / 8),      out_h = v1_h + p0 (v0_h - v1_h).
// And q = LN2(x) Wq^T enters only through q_h . k_h = LN2(x) . u_h with u_h = Wq[h*64:(h+1)*64, :]^T k_h (a D-vector per
// head and key, xattn_fold_keys_kernel below, once per forward / once per generation in the sampler).  So the whole
// D x D q-projection (38.7 GFLOP per layer at T = 32768) collapses into 12 dot products of length D per row:
//     dlt_h = LN2(x) . (u0_h - u1_h) = rstd (x . dl_h - mean sum(dl_h)) + sum(beta2 (u0_h - u1_h)),   dl_h = gamma2 (u0_h - u1_h).
// The same warp then has the new row x + out in registers and produces norm3 of it for the MLP, so norm2, the q GEMM with
// its 2-key epilogue and norm3 (3 kernels, 23 + 42 + 23 us, 550 MB) become one pass: read x, write x, write bf16 LN3(x).
// The logits are formed from the fp32 row (the GEMM path rounds LN2(x) and q's operands to bf16).
//
// Layout: one warp = 4 rows (lane l holds columns 128 j + 4 l .. + 3 of each, j < V = D / 128, so its columns of block j
// belong to head 2 j + (l >= 16)); a CTA walks a contiguous range of 4-row groups and keeps the current sample's tables in
// shared memory: dl [H][D], v1, v0 - v1, gamma3, beta3, per-head sum(dl_h) and the beta constant.  Every dl element read
// from shared memory feeds 4 rows (FMA : LDS.128 = 16 : 1); four-row reductions use a 6-shuffle transpose-reduce.
// Bound: HBM (D * 10 bytes per row) with ~17 k FMA-pipe operations per row riding along.
#include "../../include/tld_b200.h"
#include "common.h"
#include "launch.h"
#include "ptx.cuh"

namespace tld {

static constexpr float LN_EPS = 1e-5f;
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// u[r, h, k] = sum_{d < 64} key[r, h*64 + d] * Wq[h*64 + d, k]: block = (head, 16 rows of the K/V table), thread = columns k
__global__ void __launch_bounds__(256) xattn_fold_keys_kernel(const float* __restrict__ kv, long long kv_stride, int R,
                                                              const bf16* __restrict__ wq, float* __restrict__ uk,
                                                              long long uk_stride, int D) {
  __shared__ float keys[16][64];
  const int h = blockIdx.x, r0 = blockIdx.y * 16;
  for (int i = threadIdx.x; i < 16 * 64; i += 256) {
    const int r = i >> 6, d = i & 63;
    keys[r][d] = r0 + r < R ? kv[(size_t)(r0 + r) * kv_stride + h * 64 + d] : 0.f;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < D; k += 256) {
    float acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = 0.f;
    const bf16* wp = wq + (size_t)h * 64 * D + k;
#pragma unroll 4
    for (int d = 0; d < 64; ++d) {
      const float w = __bfloat162float(wp[(size_t)d * D]);
#pragma unroll
      for (int r = 0; r < 16; ++r) acc[r] = fmaf(keys[r][d], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r)
      if (r0 + r < R) uk[(size_t)(r0 + r) * uk_stride + (size_t)h * D + k] = acc[r];
  }
}

int launch_xattn_fold_keys(const float* kv, long long kv_stride, int R, const bf16* wq, float* uk, long long uk_stride, int D,
                           cudaStream_t st) {
  TLD_CHECK(kv && wq && uk && R > 0 && D % 64 == 0, "xattn_fold_keys: bad argument");
  xattn_fold_keys_kernel<<<dim3(D / 64, (R + 15) / 16), 256, 0, st>>>(kv, kv_stride, R, wq, uk, uk_stride, D);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

struct XlArgs {
  float* x;                        // [rows, D] residual stream, updated in place
  const float *g2, *b2, *g3, *b3;  // norm2 / norm3 affine
  const float *uk0, *uk1;          // folded keys [.., H, D] of this layer: noise-token rows, label rows
  long long uk0_stride, uk1_stride;
  const float *kv0, *kv1;          // K|V rows of this layer (V at + D)
  long long kv0_stride, kv1_stride;
  const int* step_ptr;             // if non-null every sample uses noise row *step_ptr
  bf16* y;                         // [rows, D] norm3(x_new)
  int rows, n_tok;
};

__device__ __forceinline__ uint32_t tf32_rna(float f) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(f));
  return u;
}

// Per-sample tables of both row kernels -> shared memory: dl_h = gamma2 (u0 - u1)_h per head, the per-head constants
// hs = sum_c dl_h[c] and hc = sum_c beta2[c] (u0 - u1)_h[c], v1 and v0 - v1.  Every thread issues ALL its global loads first
// (one memory latency per build; the per-head loop of the first version paid three in a row and idled half the warps on its
// second pass: 24 % of the kernel's samples in the ncu capture).  Item = one float4 of the [H][D / 4] table, consecutive
// threads take consecutive items, so the 32 items of a warp lie inside one head and the two sums are warp reductions into
// part[H][V][2], added up in a fixed order afterwards (deterministic).
// MODE 0: dl in fp32, [H][D] (FFMA kernel); 1: dl rounded to tf32 (cvt.rna), hs over the rounded values, heads h and h + 8
// interleaved element by element in 8 rows of 2 D + 4 floats (mma kernel); 2: dl split into tf32 upper / lower parts
// (dl_lo = second table of the same layout).
template <int V, int NT, int MODE>
__device__ __forceinline__ void xl_build_tables(const XlArgs& a, long long r0, int b, const float4* g2s, const float4* b2s, float* dl,
                                                float* dl_lo, float4* v1, float4* dv, float* hs, float* hc, float* part) {
  constexpr int D = V * 128, H = 2 * V, D4 = D / 4, ITEMS = H * D4, PER = (ITEMS + NT - 1) / NT;
  constexpr uint32_t HI_MASK = 0xffffe000u;
  const int lane = threadIdx.x & 31;
  const float4* u0 = reinterpret_cast<const float4*>(a.uk0 + r0 * a.uk0_stride);
  const float4* u1 = reinterpret_cast<const float4*>(a.uk1 + (long long)b * a.uk1_stride);
  const float4* va = reinterpret_cast<const float4*>(a.kv0 + r0 * a.kv0_stride) + D4;
  const float4* vb = reinterpret_cast<const float4*>(a.kv1 + (long long)b * a.kv1_stride) + D4;
  constexpr int PERV = (D4 + NT - 1) / NT;
  float4 p[PER], q[PER], pv[PERV], qv[PERV];
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int item = threadIdx.x + NT * k;
    if (item < ITEMS) {
      p[k] = __ldg(u0 + item);
      q[k] = __ldg(u1 + item);
    }
  }
#pragma unroll
  for (int k = 0; k < PERV; ++k) {
    const int i = threadIdx.x + NT * k;
    if (i < D4) {
      pv[k] = __ldg(va + i);
      qv[k] = __ldg(vb + i);
    }
  }
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int item = threadIdx.x + NT * k;
    if (item < ITEMS) {   // warp-uniform: ITEMS and NT are multiples of 32
      const int h = item / D4, i = item % D4;
      const float4 ga = g2s[i], be = b2s[i];
      const float4 d = make_float4(p[k].x - q[k].x, p[k].y - q[k].y, p[k].z - q[k].z, p[k].w - q[k].w);
      const float o[4] = {d.x * ga.x, d.y * ga.y, d.z * ga.z, d.w * ga.w};
      float hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (MODE == 0) {
          hi[e] = o[e];
          lo[e] = 0.f;
        } else if (MODE == 1) {
          hi[e] = __uint_as_float(tf32_rna(o[e]));
          lo[e] = 0.f;
        } else {
          hi[e] = __uint_as_float(__float_as_uint(o[e]) & HI_MASK);
          lo[e] = __uint_as_float(__float_as_uint(o[e] - hi[e]) & HI_MASK);
        }
      }
      if (MODE == 0) {
        *reinterpret_cast<float4*>(dl + h * D + 4 * i) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      } else {   // heads h and h + 8 interleaved element by element (the A fragment of the mma kernel is one LDS.128 of this)
        float* dp = dl + (h & 7) * (2 * D + 4) + 8 * i + (h >> 3);
#pragma unroll
        for (int e = 0; e < 4; ++e) dp[2 * e] = hi[e];
        if (MODE == 2) {
          float* dq = dl_lo + (h & 7) * (2 * D + 4) + 8 * i + (h >> 3);
#pragma unroll
          for (int e = 0; e < 4; ++e) dq[2 * e] = lo[e];
        }
      }
      float s = ((hi[0] + lo[0]) + (hi[1] + lo[1])) + ((hi[2] + lo[2]) + (hi[3] + lo[3]));
      float c = (d.x * be.x + d.y * be.y) + (d.z * be.z + d.w * be.w);
      s = warp_sum(s);
      c = warp_sum(c);
      if (lane == 0) {
        part[(h * V + i / 32) * 2] = s;
        part[(h * V + i / 32) * 2 + 1] = c;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < PERV; ++k) {
    const int i = threadIdx.x + NT * k;
    if (i < D4) {
      v1[i] = qv[k];
      dv[i] = make_float4(pv[k].x - qv[k].x, pv[k].y - qv[k].y, pv[k].z - qv[k].z, pv[k].w - qv[k].w);
    }
  }
  __syncthreads();
  if (threadIdx.x < H) {
    float s = 0.f, c = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      s += part[(threadIdx.x * V + j) * 2];
      c += part[(threadIdx.x * V + j) * 2 + 1];
    }
    hs[threadIdx.x] = s;
    hc[threadIdx.x] = c;
  }
  __syncthreads();
}
// groups [G0, G1) of GR rows of this CTA: an even split of the rows over the grid.  (One CTA per sample - a single table
// build per CTA - was measured too: 128 CTAs on 148 SMs carry 16 % more rows each, which costs what the saved builds gain.)
__device__ __forceinline__ void xl_cta_range(const XlArgs& a, int GR, int& G0, int& G1) {
  const int n_groups = a.rows / GR;
  G0 = int((long long)n_groups * blockIdx.x / gridDim.x);
  G1 = int((long long)n_groups * (blockIdx.x + 1) / gridDim.x);
}

// 1-D bulk copy global -> this CTA's shared memory, completion (bytes) on an mbarrier: the warp's next 4 rows (contiguous in x)
// stream in while it computes on the current ones
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// totals of R per-lane values over the warp (transpose-reduce).  R = 4: afterwards the lane holds the total of value
// 2 * bit4(lane) + bit3(lane) (6 shuffles); R = 2: of value bit4(lane) (5 shuffles)
template <int R>
__device__ __forceinline__ float reduce_rows(const float (&v)[R], int lane) {
  const bool hi4 = lane & 16, hi3 = lane & 8;
  float k;
  if (R == 4) {
    float k0 = hi4 ? v[2] : v[0], k1 = hi4 ? v[R - 1] : v[1];
    k0 += __shfl_xor_sync(0xffffffffu, hi4 ? v[0] : v[2], 16);
    k1 += __shfl_xor_sync(0xffffffffu, hi4 ? v[1] : v[R - 1], 16);
    k = hi3 ? k1 : k0;
    k += __shfl_xor_sync(0xffffffffu, hi3 ? k0 : k1, 8);
  } else {
    k = hi4 ? v[1] : v[0];
    k += __shfl_xor_sync(0xffffffffu, hi4 ? v[0] : v[1], 16);
    k += __shfl_xor_sync(0xffffffffu, k, 8);
  }
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;
}
// value of the lane group that owns row r (see reduce_rows) -> every lane
template <int R>
__device__ __forceinline__ void bcast_rows(float mine, int lane, float (&out)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r)
    out[r] = __shfl_sync(0xffffffffu, mine, R == 4 ? (r >> 1) * 16 + (r & 1) * 8 + (lane & 7) : r * 16 + (lane & 15));
}
// Two values per owner group in ONE shuffle: the owner lanes with bit 2 clear offer `lo`, those with bit 2 set offer `hi`;
// lanes 0..15 fetch row r's `lo`, lanes 16..31 its `hi` (every lane of an owner group holds both).
template <int R>
__device__ __forceinline__ void bcast_rows_split(float lo, float hi, int lane, float (&out)[R]) {
  const float mine = (lane & 4) ? hi : lo;
  const int sub = ((lane >> 4) & 1) * 4 + (lane & 3);
#pragma unroll
  for (int r = 0; r < R; ++r) out[r] = __shfl_sync(0xffffffffu, mine, R == 4 ? (r >> 1) * 16 + (r & 1) * 8 + sub : r * 16 + sub);
}

// R rows per warp, 32 / R warps.  R = 4: 8 warps x 255 registers, every dl element read from shared memory feeds 4 rows.
// (12 warps x 4 rows at 168 registers spilled 250 B at D = 768 and measured slower: 75.5 against 68.6 us.)
// R = 2: 16 warps x 128 registers - twice the shared-memory reads and 1.7 x the shuffles per row, but four warps per
// scheduler instead of two: the R = 4 kernel issues on 33 % of the cycles, each warp one instruction every ~4 cycles
// (fixed-latency dependencies and LDS waits it has nobody to hide behind).
// CT = 2: the same 32 rows in flight per SM as two independent CTAs of half the warps - the table builds and their
// barriers stall only half the SM, and the two CTAs drift apart, so one CTA's loads / stores overlap the other's arithmetic.
__host__ __device__ constexpr int xl_warps(int R, int CT) { return 32 / (R * CT); }
template <int V, int R, int CT>
__global__ void __launch_bounds__(xl_warps(R, CT) * 32, CT) ln_xattn_ln_kernel(XlArgs a) {
  constexpr int NW = xl_warps(R, CT), NT = NW * 32;
  constexpr int D = V * 128, H = 2 * V, D4 = D / 4;
  extern __shared__ float4 xl_smem[];
  float4* dl = xl_smem;            // [H][D4]
  float4* v1 = dl + H * D4;        // [D4]
  float4* dv = v1 + D4;
  float4* g3 = dv + D4;
  float4* b3 = g3 + D4;
  float4* rowbuf = b3 + D4;                        // [NW warps][R rows][D4]: the warp's next rows, filled by a bulk copy
  float4* g2s = rowbuf + NW * R * D4;              // [D4] gamma2, beta2 (table builds)
  float4* b2s = g2s + D4;
  float* hs = reinterpret_cast<float*>(b2s + D4);  // [H] sum_c dl_h[c]
  float* hc = hs + H;                              // [H] sum_c beta2[c] (u0 - u1)_h[c]
  float* part = hc + H;                            // [H][V][2] partial sums of a table build
  uint64_t* row_bar = reinterpret_cast<uint64_t*>(part + H * V * 2);   // [NW]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work unit = group of R consecutive rows (one warp); a CTA owns a contiguous range of groups and walks it sample by
  // sample (segment = the part of a sample inside the range): warp w takes groups seg + w, seg + w + NW, ...
  const int gps = a.n_tok / R;   // groups per sample
  int G0, G1;
  xl_cta_range(a, R, G0, G1);
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < D4; i += NT) {
    g3[i] = __ldg(reinterpret_cast<const float4*>(a.g3) + i);
    b3[i] = __ldg(reinterpret_cast<const float4*>(a.b3) + i);
    g2s[i] = __ldg(reinterpret_cast<const float4*>(a.g2) + i);
    b2s[i] = __ldg(reinterpret_cast<const float4*>(a.b2) + i);
  }
  if (threadIdx.x < NW) mbar_init(&row_bar[threadIdx.x], 1);
  fence_mbar_init();
  __syncthreads();
  pdl_wait();
  const long long step_row = a.step_ptr ? (long long)(*a.step_ptr) : 0;   // one dependent load per kernel, not per table build
  float4* my_buf = rowbuf + warp * R * D4;
  auto seg_end_of = [&](int g) { const int e = (g / gps + 1) * gps; return e < G1 ? e : G1; };
  auto first_from = [&](int s) {   // first group >= segment start s that belongs to this warp, or -1
    while (s < G1) {
      const int e = seg_end_of(s);
      if (s + warp < e) return s + warp;
      s = e;
    }
    return -1;
  };
  auto next_of = [&](int g, int seg_end) { return g + NW < seg_end ? g + NW : first_from(seg_end); };
  auto prefetch_rows = [&](int g) {   // lane 0: the R rows of group g (R D floats, contiguous in x)
    mbar_expect_tx(&row_bar[warp], R * D * 4);
    bulk_load_1d(smem_u32(my_buf), a.x + (size_t)g * R * D, R * D * 4, &row_bar[warp]);
  };
  {
    const int g = first_from(G0);
    if (lane == 0 && g >= 0) prefetch_rows(g);
  }
  uint32_t row_phase = 0;
  const float scale = 0.125f;   // 1 / sqrt(head_dim)
  for (int seg = G0; seg < G1;) {
    const int seg_end = seg_end_of(seg);
    const int b = seg / gps;
    {   // the sample's tables
      __syncthreads();  // every warp is done with the previous sample's tables
      const long long r0 = a.step_ptr ? step_row : (long long)b;
      xl_build_tables<V, NT, 0>(a, r0, b, g2s, b2s, reinterpret_cast<float*>(dl), nullptr, v1, dv, hs, hc, part);
    }
    // ---- this warp's groups of the segment
    for (int gw = seg + warp; gw < seg_end; gw += NW) {
    const int row = gw * R;
    float4 xv[R][V];
    float acc[R];
    mbar_wait(&row_bar[warp], row_phase);
    row_phase ^= 1;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        xv[r][j] = my_buf[r * D4 + lane + 32 * j];
        s += (xv[r][j].x + xv[r][j].y) + (xv[r][j].z + xv[r][j].w);
      }
      acc[r] = s;
    }
    __syncwarp();   // every lane has its rows in registers: the buffer may be refilled
    {
      const int gn = next_of(gw, seg_end);
      if (lane == 0 && gn >= 0) prefetch_rows(gn);
    }
    const float mean_mine = reduce_rows<R>(acc, lane) * (1.f / D);
    float mean[R];
    bcast_rows<R>(mean_mine, lane, mean);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float e0 = xv[r][j].x - mean[r], e1 = xv[r][j].y - mean[r], e2 = xv[r][j].z - mean[r], e3 = xv[r][j].w - mean[r];
        q += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
      }
      acc[r] = q;
    }
    const float rstd_mine = rsqrtf(reduce_rows<R>(acc, lane) * (1.f / D) + LN_EPS);
    // ---- per head: dlt = LN2(x) . (u0 - u1), p0 = sigmoid(dlt / 8); the lane keeps p0 of the heads its columns belong to.
    // Heads go in batches of HB: first the dots of the batch, then its HB transpose-reductions side by side (independent
    // shuffle chains: one round trip of latency per stage for the whole batch), then the sigmoids with an approximate
    // reciprocal.  Head by head, every head ended in six dependent shuffle round trips and an IEEE division whose slow-path
    // branch (BSSY / BSYNC) kept ptxas from overlapping anything with the next head: ~250 exposed cycles per head.
    constexpr int HB = H % 6 == 0 ? 6 : (H % 4 == 0 ? 4 : 2);
    float p0h[H];   // p0 of the row this lane owns after reduce_rows, for every head; broadcast when the columns are updated
#pragma unroll
    for (int h0 = 0; h0 < H; h0 += HB) {
      float part[HB][R];
#pragma unroll
      for (int hh = 0; hh < HB; ++hh) {
        const int h = h0 + hh;
        float2 a2[R];   // packed fp32 FMA (FFMA2): half the issue slots of the dot products
#pragma unroll
        for (int r = 0; r < R; ++r) a2[r] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float4 d = dl[h * D4 + lane + 32 * j];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            a2[r] = ffma2(make_float2(xv[r][j].x, xv[r][j].y), make_float2(d.x, d.y), a2[r]);
            a2[r] = ffma2(make_float2(xv[r][j].z, xv[r][j].w), make_float2(d.z, d.w), a2[r]);
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) part[hh][r] = a2[r].x + a2[r].y;
      }
      float tot[HB];
#pragma unroll
      for (int hh = 0; hh < HB; ++hh) tot[hh] = reduce_rows<R>(part[hh], lane);
#pragma unroll
      for (int hh = 0; hh < HB; ++hh) {
        const float dlt = rstd_mine * (tot[hh] - mean_mine * hs[h0 + hh]) + hc[h0 + hh];
        p0h[h0 + hh] = __fdividef(1.f, 1.f + __expf(-dlt * scale));
      }
    }
    // ---- x_new = x + v1 + p0 (v0 - v1); norm3 statistics
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float4 q = v1[lane + 32 * j], d = dv[lane + 32 * j];
      // columns 128 j + 4 lane .. lie in head 2 j + (lane >= 16): fetch that head's p0 of every row from the lanes that own the row
      float pab[R];   // lanes 0..15: p0 of head 2 j, lanes 16..31: of head 2 j + 1, for every row (one shuffle per row)
      bcast_rows_split<R>(p0h[2 * j], p0h[2 * j + 1], lane, pab);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float pr = pab[r];
        xv[r][j].x += fmaf(pr, d.x, q.x);
        xv[r][j].y += fmaf(pr, d.y, q.y);
        xv[r][j].z += fmaf(pr, d.z, q.z);
        xv[r][j].w += fmaf(pr, d.w, q.w);
        acc[r] += (xv[r][j].x + xv[r][j].y) + (xv[r][j].z + xv[r][j].w);
        // the new residual row leaves here, ahead of the norm3 reductions: 48 stores per lane in one burst at the end of the
        // group was 19 % of the kernel's samples in lg_throttle
        reinterpret_cast<float4*>(a.x + (size_t)(row + r) * D)[lane + 32 * j] = xv[r][j];
      }
    }
    bcast_rows<R>(reduce_rows<R>(acc, lane) * (1.f / D), lane, mean);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float e0 = xv[r][j].x - mean[r], e1 = xv[r][j].y - mean[r], e2 = xv[r][j].z - mean[r], e3 = xv[r][j].w - mean[r];
        q += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
      }
      acc[r] = q;
    }
    float rstd[R];
    bcast_rows<R>(rsqrtf(reduce_rows<R>(acc, lane) * (1.f / D) + LN_EPS), lane, rstd);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      uint2* yr = reinterpret_cast<uint2*>(a.y + (size_t)(row + r) * D);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float4 g = g3[lane + 32 * j], be = b3[lane + 32 * j];
        uint2 o;
        o.x = pack_bf16x2((xv[r][j].x - mean[r]) * rstd[r] * g.x + be.x, (xv[r][j].y - mean[r]) * rstd[r] * g.y + be.y);
        o.y = pack_bf16x2((xv[r][j].z - mean[r]) * rstd[r] * g.z + be.z, (xv[r][j].w - mean[r]) * rstd[r] * g.w + be.w);
        yr[lane + 32 * j] = o;
      }
    }
    }   // groups of this warp
    seg = seg_end;
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// The same pass with the H dot products of a row on the tensor pipe (mma.sync m16n8k8 tf32, fp32 accumulate).
//
// The FFMA kernel above spends ~17 k FMA-pipe operations per row on dl_h . x (H = 12 dots of length 768) and is bound by
// instruction issue at 8 warps per SM (71 us against 38 us of HBM time).  Here the dots of 8 rows are ONE 16 x 8 x D product
//     T[h, r] = sum_c dl[h, c] x[r, c]        A = dl (heads in the M rows, H <= 16), B = x^T (the 8 rows in the N columns)
// so a warp owns 8 consecutive rows and lane (g = lane / 4, t = lane % 4) holds a QUARTER of row g in registers: columns
// 16 i + 4 t .. + 3 for i < D / 16 - exactly the B fragment of that lane when chunk i is fed as two k-steps
// (k = t <-> column 16 i + 4 t + 2 s, k = t + 4 <-> column 16 i + 4 t + 2 s + 1, s = 0 / 1; any k order is valid as long as
// A uses the same one; the table interleaves heads g and g + 8 element by element, so the A fragment of a k-step is ONE
// LDS.128, conflict-free with a row stride of 2 D + 4 floats).  Row statistics are quad reductions (2 shuffles); the logits come back in the C
// fragment (head g / g + 8, rows 2 t / 2 t + 1), go through a [8][H] per-warp scratch as p0 and every lane of the row's quad
// picks the heads of its columns up from there.
// Precision (template PREC): x is split x = hi + lo with hi = the upper 19 bits (a tf32 number) and lo = x - hi (exact), two
// products -> x enters with ~2^-20 relative error; dl is rounded to tf32 ONCE per sample (cvt.rna, 2^-12) and sum_c dl_h[c]
// is taken over the rounded values, so the kernel computes LN2(x) . dl' exactly for a dl' within 2^-12 of dl (PREC 1, the
// default: the logits differ from the FFMA kernel's by ~1e-4 relative, the q GEMM path this replaces rounds both operands
// to bf16 = 2^-9).  PREC 2 also splits dl (three products, fp32-level logits), PREC 0 feeds the fp32 bits of x as they are
// (the tensor core reads the upper 19: x truncated to tf32, one product, no extra instructions).
// No shared-memory staging of the rows: the 48 LDG.128 of a lane are all in flight at once (24 KB per warp), the next group
// of the warp is pulled into L2 by one bulk prefetch while the current one is processed.
// MEASURED (B200, T = 32768, D = 768; profiles/r02_xattn_row_kernels.txt): correct in all three precisions, but NOT faster
// than the FFMA kernel - 78 us against 72 us.  The dots were never the bound: both kernels sit at 25-34 % issue utilisation
// with 8 warps per SM; this one trades the FMA work for a quad-per-row layout whose every LDG / STG touches 8 rows x 64 B
// (twice the L1 wavefronts of the row-per-warp layout) and exposes the row loads (24 % of the samples in long_scoreboard,
// 22 % in the store phase).  Loading the next group chunk by chunk behind the stores of the current one made it worse
// (82 us).  Kept as an option (xattn_mma = 1 / 2 / 3), off by default.
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                                uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// Packed fp32 pairs held as 64-bit values from the load to the store: the row lives in 2 x D / 16 of them per lane, every
// element-wise step is one FADD2 / FFMA2 on a pair IN PLACE.  (Going through float4 / float2 variables made ptxas gather
// the pairs into fresh quads for every 128-bit store and scatter them after every load: 880 MOVs per 8-row group.)
using u64 = unsigned long long;
__device__ __forceinline__ u64 f2_pack(float a, float b) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ void f2_unpack_bits(u64 v, uint32_t& a, uint32_t& b) { asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v)); }
__device__ __forceinline__ u64 f2_add(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 f2_fma(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float f2_hsum(u64 a, u64 b) {   // (a.x + a.y) + (b.x + b.y)
  float ax, ay, bx, by;
  f2_unpack(a, ax, ay);
  f2_unpack(b, bx, by);
  return (ax + ay) + (bx + by);
}

static constexpr int XM_WARPS = 8;     // 8 warps x <= 255 registers, one CTA per SM
static constexpr int XM_PSTRIDE = 20;  // p0 scratch row stride (floats): writes (40 t + g) and reads (20 g + h) hit distinct banks
template <int V, int PREC>
__global__ void __launch_bounds__(XM_WARPS * 32, 1) ln_xattn_ln_mma_kernel(XlArgs a) {
  constexpr int NW = XM_WARPS, NT = NW * 32;
  constexpr int D = V * 128, H = 2 * V, D4 = D / 4, NC = D / 16, RS = 2 * D + 4, TBL = 8 * RS;
  constexpr uint32_t HI_MASK = 0xffffe000u;   // sign, exponent, 10 mantissa bits: the bits a tf32 operand keeps
  extern __shared__ float4 xl_smem[];
  // [8][RS] tf32 bit patterns, row gp = heads gp and gp + 8 interleaved: {dl[gp][c], dl[gp + 8][c]} for c < D (PREC 2: upper
  // part).  RS = 2 D + 4 floats: the eight 16-byte reads of a quarter warp (2 rows x 4 t) tile the 32 banks.
  float* dl = reinterpret_cast<float*>(xl_smem);
  float* dl_lo = dl + TBL;                                // PREC 2 only
  float4* v1 = reinterpret_cast<float4*>(dl_lo + (PREC == 2 ? TBL : 0));   // [D4]
  float4* dv = v1 + D4;
  float4* g3 = dv + D4;
  float4* b3 = g3 + D4;
  float4* g2s = b3 + D4;                                  // [D4] gamma2, beta2 (table builds)
  float4* b2s = g2s + D4;
  float* hs = reinterpret_cast<float*>(b2s + D4);         // [16] sum_c dl'_h[c]
  float* hc = hs + 16;                                    // [16] sum_c beta2[c] (u0 - u1)_h[c]
  float* part = hc + 16;                                  // [H][V][2] partial sums of a table build
  float* pw = part + H * V * 2;                           // [NW][8][XM_PSTRIDE] p0 of the warp's 8 rows
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // work unit = group of 8 consecutive rows (one warp); CTA ranges and the walk over samples as in the FFMA kernel
  const int gps = a.n_tok / 8;
  int G0, G1;
  xl_cta_range(a, 8, G0, G1);
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < D4; i += NT) {
    g3[i] = __ldg(reinterpret_cast<const float4*>(a.g3) + i);
    b3[i] = __ldg(reinterpret_cast<const float4*>(a.b3) + i);
    g2s[i] = __ldg(reinterpret_cast<const float4*>(a.g2) + i);
    b2s[i] = __ldg(reinterpret_cast<const float4*>(a.b2) + i);
  }
  if (H < 16)   // the slots of heads >= H (M rows of the MMA without a head) are never written by a table build
    for (int i = threadIdx.x; i < (PREC == 2 ? 2 : 1) * TBL; i += NT) dl[i] = 0.f;
  __syncthreads();
  pdl_wait();
  auto seg_end_of = [&](int gr) { const int e = (gr / gps + 1) * gps; return e < G1 ? e : G1; };
  auto first_from = [&](int s) {   // first group >= segment start s that belongs to this warp, or -1
    while (s < G1) {
      const int e = seg_end_of(s);
      if (s + warp < e) return s + warp;
      s = e;
    }
    return -1;
  };
  float* pww = pw + warp * 8 * XM_PSTRIDE;
  const float scale = 0.125f;   // 1 / sqrt(head_dim)
  for (int seg = G0; seg < G1;) {
    const int seg_end = seg_end_of(seg);
    const int b = seg / gps;
    if (seg == G0 && lane == 0 && seg + warp < seg_end) prefetch_l2_bulk(a.x + (size_t)(seg + warp) * 8 * D, 8 * D * 4);
    {   // the sample's tables
      __syncthreads();   // every warp is done with the previous sample's tables
      const long long r0 = a.step_ptr ? (long long)(*a.step_ptr) : (long long)b;
      xl_build_tables<V, NT, PREC == 2 ? 2 : 1>(a, r0, b, g2s, b2s, dl, dl_lo, v1, dv, hs, hc, part);
    }
    for (int gw = seg + warp; gw < seg_end; gw += NW) {
      // ---- the lane's quarter of row g of the group: NC x 16 bytes, all loads in flight together
      float* xr = a.x + ((size_t)gw * 8 + g) * D + 4 * t;
      u64 xa[NC], xb[NC];   // columns 16 i + 4 t, + 1 | + 2, + 3
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(xr + 16 * i);
        xa[i] = v.x;
        xb[i] = v.y;
      }
      {   // the warp's next group -> L2 while this one is processed
        const int gn = gw + NW < seg_end ? gw + NW : first_from(seg_end);
        if (lane == 0 && gn >= 0) prefetch_l2_bulk(a.x + (size_t)gn * 8 * D, 8 * D * 4);
      }
      // ---- norm2 statistics of row g (quad = the 4 lanes of the row), one pass: sum and sum of squares
      // (the element-wise work, not the dots, is 3/4 of the instructions of this kernel's unpacked version)
      u64 s2a = 0, s2b = 0, q2a = 0, q2b = 0;
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        s2a = f2_add(s2a, xa[i]);
        s2b = f2_add(s2b, xb[i]);
        q2a = f2_fma(xa[i], xa[i], q2a);
        q2b = f2_fma(xb[i], xb[i], q2b);
      }
      const float mean = quad_sum(f2_hsum(s2a, s2b)) * (1.f / D);
      const float rstd = rsqrtf(fmaxf(quad_sum(f2_hsum(q2a, q2b)) * (1.f / D) - mean * mean, 0.f) + LN_EPS);
      // ---- T[h, r] = dl_h . x_r on the tensor pipe: independent accumulator chains per product and k-step parity
      constexpr int NP = PREC == 0 ? 1 : PREC == 1 ? 2 : 3;
      float acc[2 * NP][4];
#pragma unroll
      for (int c = 0; c < 2 * NP; ++c)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[c][e] = 0.f;
      const float* dlp = dl + g * RS + 8 * t;   // {dl[g][c], dl[g + 8][c], dl[g][c + 1], dl[g + 8][c + 1]} = the A fragment of one k-step
      const u64 hi_mask2 = (u64(HI_MASK) << 32) | HI_MASK, neg1 = f2_pack(-1.f, -1.f);
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const uint4 A0 = *reinterpret_cast<const uint4*>(dlp + 32 * i), A1 = *reinterpret_cast<const uint4*>(dlp + 32 * i + 4);
        uint32_t xh[4], xl[4];
        if (PREC == 0) {   // the fp32 bits as they are: the tensor core reads the upper 19 (x truncated to tf32)
          f2_unpack_bits(xa[i], xh[0], xh[1]);
          f2_unpack_bits(xb[i], xh[2], xh[3]);
          xl[0] = xl[1] = xl[2] = xl[3] = 0;
        } else {           // xh = the upper 19 bits (a tf32 number), xl = x - xh exactly (one packed FMA per pair)
          const u64 ha2 = xa[i] & hi_mask2, hb2 = xb[i] & hi_mask2;
          f2_unpack_bits(ha2, xh[0], xh[1]);
          f2_unpack_bits(hb2, xh[2], xh[3]);
          f2_unpack_bits(f2_fma(ha2, neg1, xa[i]), xl[0], xl[1]);
          f2_unpack_bits(f2_fma(hb2, neg1, xb[i]), xl[2], xl[3]);
        }
        // k-step ss: k = t <-> column 16 i + 4 t + 2 ss, k = t + 4 <-> the next column
        mma_tf32_16x8x8(acc[0], A0.x, A0.y, A0.z, A0.w, xh[0], xh[1]);
        mma_tf32_16x8x8(acc[1], A1.x, A1.y, A1.z, A1.w, xh[2], xh[3]);
        if (PREC >= 1) {
          mma_tf32_16x8x8(acc[2], A0.x, A0.y, A0.z, A0.w, xl[0], xl[1]);
          mma_tf32_16x8x8(acc[3], A1.x, A1.y, A1.z, A1.w, xl[2], xl[3]);
        }
        if (PREC == 2) {
          const uint4 L0 = *reinterpret_cast<const uint4*>(dlp + TBL + 32 * i), L1 = *reinterpret_cast<const uint4*>(dlp + TBL + 32 * i + 4);
          mma_tf32_16x8x8(acc[4], L0.x, L0.y, L0.z, L0.w, xh[0], xh[1]);
          mma_tf32_16x8x8(acc[5], L1.x, L1.y, L1.z, L1.w, xh[2], xh[3]);
        }
      }
      float tot[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = acc[0][e] + acc[1][e];
        if (PREC >= 1) v += acc[2][e] + acc[3][e];
        if (PREC == 2) v += acc[4][e] + acc[5][e];
        tot[e] = v;
      }
      // ---- the lane holds T[g][2 t], T[g][2 t + 1], T[g + 8][2 t], T[g + 8][2 t + 1]: p0 = sigmoid(LN2 logit difference / 8) -> scratch
      {
        const float mean_a = __shfl_sync(0xffffffffu, mean, 8 * t), mean_b = __shfl_sync(0xffffffffu, mean, 8 * t + 4);
        const float rstd_a = __shfl_sync(0xffffffffu, rstd, 8 * t), rstd_b = __shfl_sync(0xffffffffu, rstd, 8 * t + 4);
        __syncwarp();   // the previous group's readers of the scratch are done
        if (g < H) {
          const float hsv = hs[g], hcv = hc[g];
          pww[(2 * t) * XM_PSTRIDE + g] = __fdividef(1.f, 1.f + __expf(-(rstd_a * (tot[0] - mean_a * hsv) + hcv) * scale));
          pww[(2 * t + 1) * XM_PSTRIDE + g] = __fdividef(1.f, 1.f + __expf(-(rstd_b * (tot[1] - mean_b * hsv) + hcv) * scale));
        }
        if (g + 8 < H) {
          const float hsv = hs[g + 8], hcv = hc[g + 8];
          pww[(2 * t) * XM_PSTRIDE + g + 8] = __fdividef(1.f, 1.f + __expf(-(rstd_a * (tot[2] - mean_a * hsv) + hcv) * scale));
          pww[(2 * t + 1) * XM_PSTRIDE + g + 8] = __fdividef(1.f, 1.f + __expf(-(rstd_b * (tot[3] - mean_b * hsv) + hcv) * scale));
        }
        __syncwarp();
      }
      // ---- x_new = x + v1 + p0 (v0 - v1): chunk i lies in head i / 4; the new row leaves at once; norm3 statistics (one pass)
      s2a = s2b = q2a = q2b = 0;
      u64 pp = 0;
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        if (i % 4 == 0) {
          const float pr = pww[g * XM_PSTRIDE + i / 4];
          pp = f2_pack(pr, pr);
        }
        const ulonglong2 qv = *reinterpret_cast<const ulonglong2*>(v1 + 4 * i + t), d = *reinterpret_cast<const ulonglong2*>(dv + 4 * i + t);
        xa[i] = f2_add(xa[i], f2_fma(pp, d.x, qv.x));
        xb[i] = f2_add(xb[i], f2_fma(pp, d.y, qv.y));
        *reinterpret_cast<ulonglong2*>(xr + 16 * i) = make_ulonglong2(xa[i], xb[i]);
        s2a = f2_add(s2a, xa[i]);
        s2b = f2_add(s2b, xb[i]);
        q2a = f2_fma(xa[i], xa[i], q2a);
        q2b = f2_fma(xb[i], xb[i], q2b);
      }
      const float mean3 = quad_sum(f2_hsum(s2a, s2b)) * (1.f / D);
      const float rstd3 = rsqrtf(fmaxf(quad_sum(f2_hsum(q2a, q2b)) * (1.f / D) - mean3 * mean3, 0.f) + LN_EPS);
      const u64 rs2 = f2_pack(rstd3, rstd3), nm2 = f2_pack(-mean3 * rstd3, -mean3 * rstd3);
      bf16* yr = a.y + ((size_t)gw * 8 + g) * D + 4 * t;
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const ulonglong2 ga = *reinterpret_cast<const ulonglong2*>(g3 + 4 * i + t), be = *reinterpret_cast<const ulonglong2*>(b3 + 4 * i + t);
        float y0, y1, y2, y3;
        f2_unpack(f2_fma(f2_fma(xa[i], rs2, nm2), ga.x, be.x), y0, y1);
        f2_unpack(f2_fma(f2_fma(xb[i], rs2, nm2), ga.y, be.y), y2, y3);
        uint2 o;
        o.x = pack_bf16x2(y0, y1);
        o.y = pack_bf16x2(y2, y3);
        *reinterpret_cast<uint2*>(yr + 16 * i) = o;
      }
    }   // groups of this warp
    seg = seg_end;
  }
}

static size_t xm_smem_bytes(int D, int prec) {   // dl table(s) + v1, dv, gamma3, beta3 + per-head constants + p0 scratch
  const int H = D / 64;
  return (size_t)(prec == 2 ? 2 : 1) * 8 * (2 * D + 4) * 4 + 6 * (size_t)D * 4 + (32 + 2 * (size_t)H * (D / 128)) * 4 +
         (size_t)XM_WARPS * 8 * XM_PSTRIDE * 4;
}

static int g_xattn_mma = 0;   // tld_set_option("xattn_mma", ...): 0 = FFMA kernel (default, faster), 1 / 2 / 3 = tensor-pipe kernel with PREC 0 / 1 / 2
void set_xattn_mma(int v) { g_xattn_mma = v; }
static bool xm_supported(int D, int n_tok) { return D % 128 == 0 && D >= 128 && D <= 768 && n_tok % 8 == 0; }

static size_t xl_smem_bytes(int D, int ctas) {   // tables + 32 / ctas rows of prefetch buffer (R rows x warps) + per-head constants + mbarriers
  const int H = D / 64, V = D / 128;
  return (size_t)H * D * 4 + 6 * (size_t)D * 4 + (32 / ctas) * (size_t)D * 4 + (2 * (size_t)H + 2 * (size_t)H * V) * 4 + 16 * 8 + 64;
}

static int g_xattn_ctas = 1;   // tld_set_option("xattn_ctas", ...): CTAs per SM of the FFMA kernel (1 or 2; embed_dim > 768 always runs 1)
void set_xattn_ctas(int v) { g_xattn_ctas = v; }
static int g_xattn_rows = 4;   // tld_set_option("xattn_rows", ...): rows per warp of the FFMA kernel, 4 (8 warps) or 2 (16 warps)
void set_xattn_rows(int v) { g_xattn_rows = v; }

bool ln_xattn_ln_supported(int D, int n_tok) { return D % 128 == 0 && D >= 128 && D <= 1024 && n_tok % 4 == 0; }   // (2 rows per warp: n_tok % 2, implied)

int launch_ln_xattn_ln(float* x, const float* g2, const float* b2, const float* g3, const float* b3, const float* uk0,
                       long long uk0_stride, const float* uk1, long long uk1_stride, const float* kv0, long long kv0_stride,
                       const float* kv1, long long kv1_stride, const int* step_ptr, bf16* y, int rows, int n_tok, int D,
                       cudaStream_t st) {
  TLD_CHECK(ln_xattn_ln_supported(D, n_tok), "ln_xattn_ln: needs embed_dim % 128 == 0 (<= 1024) and tokens per sample % 4 == 0");
  TLD_CHECK(rows > 0 && rows % n_tok == 0, "ln_xattn_ln: rows must be whole samples");
  TLD_CHECK(x && g2 && b2 && g3 && b3 && uk0 && uk1 && kv0 && kv1 && y, "ln_xattn_ln: null argument");
  TLD_CHECK(uk0_stride % 4 == 0 && uk1_stride % 4 == 0 && kv0_stride % 4 == 0 && kv1_stride % 4 == 0 &&
                ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(uk0) | reinterpret_cast<uintptr_t>(uk1) |
                  reinterpret_cast<uintptr_t>(kv0) | reinterpret_cast<uintptr_t>(kv1) | reinterpret_cast<uintptr_t>(y) |
                  reinterpret_cast<uintptr_t>(g2) | reinterpret_cast<uintptr_t>(b2) | reinterpret_cast<uintptr_t>(g3) |
                  reinterpret_cast<uintptr_t>(b3)) & 15) == 0,
            "ln_xattn_ln: operands must be 16-byte aligned");
  XlArgs a{x, g2, b2, g3, b3, uk0, uk1, uk0_stride, uk1_stride, kv0, kv1, kv0_stride, kv1_stride, step_ptr, y, rows, n_tok};
  if (g_xattn_mma && xm_supported(D, n_tok)) {   // dot products on the tensor pipe (8-row groups)
    const int prec = g_xattn_mma - 1, groups8 = rows / 8;
    const int want = (groups8 + XM_WARPS - 1) / XM_WARPS;
    const int grid = want < sm_count() ? want : sm_count();
    const size_t smem = xm_smem_bytes(D, prec);
#define XM_CASE(V, P)                                                                                                       \
  if (D == V * 128 && prec == P) {                                                                                          \
    static bool attr_set = false;                                                                                           \
    if (!attr_set) {                                                                                                        \
      TLD_CUDA_OK(cudaFuncSetAttribute(ln_xattn_ln_mma_kernel<V, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      attr_set = true;                                                                                                      \
    }                                                                                                                       \
    if (launch_pdl(ln_xattn_ln_mma_kernel<V, P>, dim3(grid), dim3(XM_WARPS * 32), smem, st, a)) return 1;                  \
    TLD_CUDA_OK(cudaGetLastError());                                                                                        \
    return 0;                                                                                                               \
  }
#define XM_CASES(V) XM_CASE(V, 0) XM_CASE(V, 1) XM_CASE(V, 2)
    XM_CASES(1) XM_CASES(2) XM_CASES(3) XM_CASES(4) XM_CASES(5) XM_CASES(6)
#undef XM_CASES
#undef XM_CASE
  }
  const int ctas = D <= 768 ? g_xattn_ctas : 1;   // two CTAs per SM need 2 x (tables + 16 rows) of shared memory
  const size_t smem = xl_smem_bytes(D, ctas);
#define XL_CASE(V, R, CT)                                                                                                   \
  if (D == V * 128 && g_xattn_rows == R && ctas == CT) {                                                                    \
    static int occ = 0;                                                                                                     \
    if (!occ) {                                                                                                             \
      TLD_CUDA_OK(cudaFuncSetAttribute(ln_xattn_ln_kernel<V, R, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      TLD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ln_xattn_ln_kernel<V, R, CT>, xl_warps(R, CT) * 32, smem)); \
      if (occ < 1) occ = 1;                                                                                                 \
    }                                                                                                                       \
    const int slots = occ * sm_count();                                                                                     \
    const int want = (rows / R + xl_warps(R, CT) - 1) / xl_warps(R, CT);                                                    \
    const int grid = want < slots ? want : slots;                                                                           \
    if (launch_pdl(ln_xattn_ln_kernel<V, R, CT>, dim3(grid), dim3(xl_warps(R, CT) * 32), smem, st, a)) return 1;            \
  }
#define XL_CASES(V) XL_CASE(V, 4, 1) XL_CASE(V, 2, 1) XL_CASE(V, 4, 2) XL_CASE(V, 2, 2)
  XL_CASES(1) XL_CASES(2) XL_CASES(3) XL_CASES(4) XL_CASES(5) XL_CASES(6)
  XL_CASE(7, 4, 1) XL_CASE(7, 2, 1) XL_CASE(8, 4, 1) XL_CASE(8, 2, 1)
#undef XL_CASES
#undef XL_CASE
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace tld
