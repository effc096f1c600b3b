// Self-attention over the latent tokens, fused with the residual add:
//     x[t, h*64:(h+1)*64] += softmax(q k^T / sqrt(64)) v        (transformer_blocks.py:31-48,57-59,136)
// qkv is the bf16 output of the qkv GEMM, [T, 3D] = (q | k | v), head_dim fixed at 64 (heads = D/64).
// There is no output projection in the reference, so the result goes straight into the fp32 residual stream.
//
// v1 kernel (round 1): flash-attention with warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate),
// one CTA = 64 query rows of one (sample, head), K/V streamed in 64-key chunks through a double-buffered,
// XOR-swizzled shared-memory ring filled with cp.async.  Self-attention is 5.2 % of the FLOPs at 256 px; the
// tcgen05 version (S in TMEM) is the follow-up once the GEMM path is at roofline.
#include "common.h"

namespace tld {

constexpr int ATT_BQ = 64;
constexpr int ATT_BKV = 64;
constexpr int ATT_HD = 64;
constexpr int ATT_THREADS = 128;

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// tile [64 rows][64 bf16] = 128 B per row, 16-byte chunk c of row r stored at chunk (c ^ (r & 7))
__device__ __forceinline__ bf16* tile_ptr(bf16* base, int r, int chunk) {
  return base + r * ATT_HD + ((chunk ^ (r & 7)) << 3);
}

__device__ __forceinline__ void load_tile_async(bf16* smem_tile, const bf16* gsrc, long long row_stride, int tid) {
  // 64 rows x 8 chunks = 512 x 16 B, 128 threads -> 4 each
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * ATT_THREADS;
    const int r = idx >> 3, c = idx & 7;
    cp_async16(tile_ptr(smem_tile, r, c), gsrc + (long long)r * row_stride + c * 8);
  }
}

__global__ void __launch_bounds__(ATT_THREADS) self_attention_kernel(const bf16* __restrict__ qkv,
                                                                     float* __restrict__ x, int n_tok, int D) {
  __shared__ __align__(128) bf16 sQ[ATT_BQ * ATT_HD];
  __shared__ __align__(128) bf16 sK[2][ATT_BKV * ATT_HD];
  __shared__ __align__(128) bf16 sV[2][ATT_BKV * ATT_HD];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const long long ld = 3LL * D;
  const bf16* qbase = qkv + ((long long)b * n_tok + (long long)qt * ATT_BQ) * ld + head * ATT_HD;
  const bf16* kbase = qkv + (long long)b * n_tok * ld + D + head * ATT_HD;
  const bf16* vbase = kbase + D;
  const int n_chunks = n_tok / ATT_BKV;

  load_tile_async(sQ, qbase, ld, tid);
  load_tile_async(sK[0], kbase, ld, tid);
  load_tile_async(sV[0], vbase, ld, tid);
  cp_async_commit();

  // Q fragments (A operand, 16 rows x 64 cols per warp) are loaded once
  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float sl2 = 0.125f * 1.4426950408889634f;  // softmax scale * log2(e)

  for (int ck = 0; ck < n_chunks; ++ck) {
    const int buf = ck & 1;
    if (ck + 1 < n_chunks) {
      load_tile_async(sK[buf ^ 1], kbase + (long long)(ck + 1) * ATT_BKV * ld, ld, tid);
      load_tile_async(sV[buf ^ 1], vbase + (long long)(ck + 1) * ATT_BKV * ld, ld, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (ck == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int chunk = ks * 2 + (lane >> 4);
        ldsm_x4(qf[ks], tile_ptr(sQ, r, chunk));
      }
    }
    // S = Q K^T : 16 x 64 per warp
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-key n-tiles
        uint32_t kf[4];
        const int key = np * 16 + (lane & 7) + (lane >> 4) * 8;
        const int chunk = ks * 2 + ((lane >> 3) & 1);
        ldsm_x4(kf, tile_ptr(sK[buf], key, chunk));
        mma_bf16(s[2 * np], qf[ks], kf[0], kf[1]);
        mma_bf16(s[2 * np + 1], qf[ks], kf[2], kf[3]);
      }
    }
    // online softmax; this thread owns rows g=lane/4 (c0,c1) and g+8 (c2,c3)
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
    float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      corr[r] = exp2f((m_run[r] - mx[r]) * sl2);
      m_run[r] = mx[r];
    }
    uint32_t pf[4][4];  // P as A-operand fragments for the 4 k-steps (16 keys each)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float p0 = exp2f((s[i][0] - mx[0]) * sl2), p1 = exp2f((s[i][1] - mx[0]) * sl2);
      const float p2 = exp2f((s[i][2] - mx[1]) * sl2), p3 = exp2f((s[i][3] - mx[1]) * sl2);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[i >> 1][(i & 1) * 2 + 0] = pack2(p0, p1);
      pf[i >> 1][(i & 1) * 2 + 1] = pack2(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i][0] *= corr[0];
      o[i][1] *= corr[0];
      o[i][2] *= corr[1];
      o[i][3] *= corr[1];
    }
    // O += P V
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {  // pairs of 8-wide d tiles
        uint32_t vf[4];
        const int key = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int chunk = dp * 2 + (lane >> 4);
        ldsm_x4_t(vf, tile_ptr(sV[buf], key, chunk));
        mma_bf16(o[2 * dp], pf[ks], vf[0], vf[1]);
        mma_bf16(o[2 * dp + 1], pf[ks], vf[2], vf[3]);
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled two iterations later
  }

  // finalise: full row sums live across the 4 lanes of a quad
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int g = lane >> 2, t4 = lane & 3;
  const long long row0 = (long long)b * n_tok + (long long)qt * ATT_BQ + warp * 16 + g;
  float* x0 = x + row0 * D + head * ATT_HD + t4 * 2;
  float* x1 = x0 + 8LL * D;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float2 a = *reinterpret_cast<float2*>(x0 + i * 8);
    a.x += o[i][0] * inv0;
    a.y += o[i][1] * inv0;
    *reinterpret_cast<float2*>(x0 + i * 8) = a;
    float2 c = *reinterpret_cast<float2*>(x1 + i * 8);
    c.x += o[i][2] * inv1;
    c.y += o[i][3] * inv1;
    *reinterpret_cast<float2*>(x1 + i * 8) = c;
  }
}

int launch_self_attention_mma(const bf16* qkv, float* x, int B, int n_tok, int D, cudaStream_t st) {
  TLD_CHECK(D % 64 == 0, "self_attention: embed_dim must be a multiple of 64");
  TLD_CHECK(n_tok % 64 == 0, "self_attention: tokens per sample must be a multiple of 64");
  TLD_CHECK(B <= 65535, "self_attention: batch too large for gridDim.z");
  dim3 grid(n_tok / ATT_BQ, D / 64, B);
  self_attention_kernel<<<grid, ATT_THREADS, 0, st>>>(qkv, x, n_tok, D);
  TLD_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_self_attention(const bf16* qkv, float* x, int B, int n_tok, int D, cudaStream_t st, int impl) {
  if (impl == 3 || (impl == 0 && n_tok % 128 == 0)) return launch_self_attention_tc2(qkv, x, B, n_tok, D, st);
  return launch_self_attention_mma(qkv, x, B, n_tok, D, st);
}

}  // namespace tld
