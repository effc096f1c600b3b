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

// 1-D bulk copy global -> this CTA's shared memory, completion (bytes) on an mbarrier: the warp's next 4 rows (contiguous in x)
// stream in while it computes on the current ones
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// totals of 4 per-lane values over the warp: afterwards the lane holds the total of value 2 * bit4(lane) + bit3(lane)
__device__ __forceinline__ float reduce4(const float (&v)[4], int lane) {
  const bool hi4 = lane & 16, hi3 = lane & 8;
  float k0 = hi4 ? v[2] : v[0], k1 = hi4 ? v[3] : v[1];
  k0 += __shfl_xor_sync(0xffffffffu, hi4 ? v[0] : v[2], 16);
  k1 += __shfl_xor_sync(0xffffffffu, hi4 ? v[1] : v[3], 16);
  float k = hi3 ? k1 : k0;
  k += __shfl_xor_sync(0xffffffffu, hi3 ? k0 : k1, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;
}
// value of the lane group that owns row r (see reduce4) -> every lane
__device__ __forceinline__ void bcast4(float mine, int lane, float (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) out[r] = __shfl_sync(0xffffffffu, mine, (r >> 1) * 16 + (r & 1) * 8 + (lane & 7));
}

// 8 warps x 255 registers.  12 warps (168 registers: 250 B of spills at D = 768, more idle warps at the sample boundaries of a
// CTA's range) measured slower: 75.5 against 68.6 us.
__host__ __device__ constexpr int xl_warps(int V) { return 8; }
template <int V>
__global__ void __launch_bounds__(xl_warps(V) * 32, 1) ln_xattn_ln_kernel(XlArgs a) {
  constexpr int NW = xl_warps(V), NT = NW * 32;
  constexpr int D = V * 128, H = 2 * V, D4 = D / 4;
  extern __shared__ float4 xl_smem[];
  float4* dl = xl_smem;            // [H][D4]
  float4* v1 = dl + H * D4;        // [D4]
  float4* dv = v1 + D4;
  float4* g3 = dv + D4;
  float4* b3 = g3 + D4;
  float4* rowbuf = b3 + D4;                        // [NW warps][4 rows][D4]: the warp's next rows, filled by a bulk copy
  float* hs = reinterpret_cast<float*>(rowbuf + NW * 4 * D4);   // [H] sum_c dl_h[c]
  float* hc = hs + H;                              // [H] sum_c beta2[c] (u0 - u1)_h[c]
  uint64_t* row_bar = reinterpret_cast<uint64_t*>(hc + H + (H & 1));   // [NW]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work unit = group of 4 consecutive rows (one warp); a CTA owns a contiguous range of groups and walks it sample by
  // sample (segment = the part of a sample inside the range): warp w takes groups seg + w, seg + w + NW, ...
  const int n_groups = a.rows / 4, gps = a.n_tok / 4;   // groups per sample
  const int G0 = int((long long)n_groups * blockIdx.x / gridDim.x), G1 = int((long long)n_groups * (blockIdx.x + 1) / gridDim.x);
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < D4; i += NT) {
    g3[i] = __ldg(reinterpret_cast<const float4*>(a.g3) + i);
    b3[i] = __ldg(reinterpret_cast<const float4*>(a.b3) + i);
  }
  if (threadIdx.x < NW) mbar_init(&row_bar[threadIdx.x], 1);
  fence_mbar_init();
  __syncthreads();
  pdl_wait();
  float4* my_buf = rowbuf + warp * 4 * D4;
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
  auto prefetch_rows = [&](int g) {   // lane 0: the 4 rows of group g (4 D floats, contiguous in x)
    mbar_expect_tx(&row_bar[warp], 4 * D * 4);
    bulk_load_1d(smem_u32(my_buf), a.x + (size_t)g * 4 * D, 4 * D * 4, &row_bar[warp]);
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
    {   // build the sample's tables
      __syncthreads();  // every warp is done with the previous sample's tables
      const long long r0 = a.step_ptr ? (long long)(*a.step_ptr) : (long long)b;
      const float4* u0 = reinterpret_cast<const float4*>(a.uk0 + r0 * a.uk0_stride);
      const float4* u1 = reinterpret_cast<const float4*>(a.uk1 + (long long)b * a.uk1_stride);
      for (int h = warp; h < H; h += NW) {   // warp = head: dl_h and its two constants
        float s = 0.f, c = 0.f;
#pragma unroll
        for (int jj = 0; jj < V; ++jj) {   // unrolled: all 4 V loads of the head in flight at once
          const int i = lane + 32 * jj;
          const float4 p = __ldg(u0 + h * D4 + i), q = __ldg(u1 + h * D4 + i);
          const float4 g = __ldg(reinterpret_cast<const float4*>(a.g2) + i), be = __ldg(reinterpret_cast<const float4*>(a.b2) + i);
          const float4 d = make_float4(p.x - q.x, p.y - q.y, p.z - q.z, p.w - q.w);
          const float4 o = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
          dl[h * D4 + i] = o;
          s += (o.x + o.y) + (o.z + o.w);
          c += (d.x * be.x + d.y * be.y) + (d.z * be.z + d.w * be.w);
        }
        s = warp_sum(s);
        c = warp_sum(c);
        if (lane == 0) {
          hs[h] = s;
          hc[h] = c;
        }
      }
      const float4* va = reinterpret_cast<const float4*>(a.kv0 + r0 * a.kv0_stride) + D4;
      const float4* vb = reinterpret_cast<const float4*>(a.kv1 + (long long)b * a.kv1_stride) + D4;
      for (int i = threadIdx.x; i < D4; i += NT) {
        const float4 p = __ldg(va + i), q = __ldg(vb + i);
        v1[i] = q;
        dv[i] = make_float4(p.x - q.x, p.y - q.y, p.z - q.z, p.w - q.w);
      }
      __syncthreads();
    }
    // ---- this warp's groups of the segment
    for (int gw = seg + warp; gw < seg_end; gw += NW) {
    const int row = gw * 4;
    float4 xv[4][V];
    float acc[4];
    mbar_wait(&row_bar[warp], row_phase);
    row_phase ^= 1;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
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
    const float mean_mine = reduce4(acc, lane) * (1.f / D);
    float mean[4];
    bcast4(mean_mine, lane, mean);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float e0 = xv[r][j].x - mean[r], e1 = xv[r][j].y - mean[r], e2 = xv[r][j].z - mean[r], e3 = xv[r][j].w - mean[r];
        q += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
      }
      acc[r] = q;
    }
    const float rstd_mine = rsqrtf(reduce4(acc, lane) * (1.f / D) + LN_EPS);
    // ---- per head: dlt = LN2(x) . (u0 - u1), p0 = sigmoid(dlt / 8); the lane keeps p0 of the heads its columns belong to
    float p0h[H];   // p0 of row (2 bit4 + bit3)(lane) for every head; broadcast when the columns are updated
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float2 a2[4];   // packed fp32 FMA (FFMA2): half the issue slots of the dot products
#pragma unroll
      for (int r = 0; r < 4; ++r) a2[r] = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float4 d = dl[h * D4 + lane + 32 * j];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          a2[r] = ffma2(make_float2(xv[r][j].x, xv[r][j].y), make_float2(d.x, d.y), a2[r]);
          a2[r] = ffma2(make_float2(xv[r][j].z, xv[r][j].w), make_float2(d.z, d.w), a2[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = a2[r].x + a2[r].y;
      const float tot = reduce4(acc, lane);
      const float dlt = rstd_mine * (tot - mean_mine * hs[h]) + hc[h];
      p0h[h] = 1.f / (1.f + __expf(-dlt * scale));
    }
    // ---- x_new = x + v1 + p0 (v0 - v1); norm3 statistics
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r] = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float4 q = v1[lane + 32 * j], d = dv[lane + 32 * j];
      // columns 128 j + 4 lane .. lie in head 2 j + (lane >= 16): fetch that head's p0 of every row from the lanes that own the row
      float pa[4], pb[4];
      bcast4(p0h[2 * j], lane, pa);
      bcast4(p0h[2 * j + 1], lane, pb);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float pr = (lane & 16) ? pb[r] : pa[r];
        xv[r][j].x += fmaf(pr, d.x, q.x);
        xv[r][j].y += fmaf(pr, d.y, q.y);
        xv[r][j].z += fmaf(pr, d.z, q.z);
        xv[r][j].w += fmaf(pr, d.w, q.w);
        acc[r] += (xv[r][j].x + xv[r][j].y) + (xv[r][j].z + xv[r][j].w);
      }
    }
    bcast4(reduce4(acc, lane) * (1.f / D), lane, mean);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float e0 = xv[r][j].x - mean[r], e1 = xv[r][j].y - mean[r], e2 = xv[r][j].z - mean[r], e3 = xv[r][j].w - mean[r];
        q += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
      }
      acc[r] = q;
    }
    float rstd[4];
    bcast4(rsqrtf(reduce4(acc, lane) * (1.f / D) + LN_EPS), lane, rstd);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float4* xr = reinterpret_cast<float4*>(a.x + (size_t)(row + r) * D);
      uint2* yr = reinterpret_cast<uint2*>(a.y + (size_t)(row + r) * D);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        xr[lane + 32 * j] = xv[r][j];
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

static size_t xl_smem_bytes(int D) {   // tables + 4 rows of prefetch buffer per warp + per-head constants + mbarriers
  const int nw = xl_warps(D / 128);
  return (size_t)(D / 64) * D * 4 + 4 * (size_t)D * 4 + nw * 4 * (size_t)D * 4 + 2 * (size_t)(D / 64 + 1) * 4 + nw * 8 + 64;
}

bool ln_xattn_ln_supported(int D, int n_tok) { return D % 128 == 0 && D >= 128 && D <= 1024 && n_tok % 4 == 0; }

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
  const int n_groups = rows / 4;
  const size_t smem = xl_smem_bytes(D);
  switch (D / 128) {
#define XL_CASE(V)                                                                                                          \
  case V: {                                                                                                                 \
    static int occ = 0;                                                                                                     \
    if (!occ) {                                                                                                             \
      TLD_CUDA_OK(cudaFuncSetAttribute(ln_xattn_ln_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xl_smem_bytes(V * 128))); \
      TLD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ln_xattn_ln_kernel<V>, xl_warps(V) * 32, xl_smem_bytes(V * 128)));  \
      if (occ < 1) occ = 1;                                                                                                 \
    }                                                                                                                       \
    const int slots = occ * sm_count();                                                                                     \
    const int want = (n_groups + xl_warps(V) - 1) / xl_warps(V);                                                            \
    const int grid = want < slots ? want : slots;                                                                           \
    if (launch_pdl(ln_xattn_ln_kernel<V>, dim3(grid), dim3(xl_warps(V) * 32), smem, st, a)) return 1;                       \
  } break;
    XL_CASE(1) XL_CASE(2) XL_CASE(3) XL_CASE(4) XL_CASE(5) XL_CASE(6) XL_CASE(7) XL_CASE(8)
#undef XL_CASE
  }
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace tld
