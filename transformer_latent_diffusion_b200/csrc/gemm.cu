// Host side of the tcgen05 GEMM: TMA descriptor encoding, tile-shape selection, launch.
#include <cudaTypedefs.h>

#include <mutex>

#include "common.h"
#include "gemm_tcgen05.cuh"
#include "launch.h"

namespace tld {

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static int resolve_encode() {
  static std::once_flag once;
  static int status = 0;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || fn == nullptr) {
      status = 1;
      return;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  });
  return status;
}

// 2-D row-major tensor [rows, cols] (cols contiguous, leading dimension ld elements) of bf16 or fp32,
// box = [box_rows, 128 bytes of columns], 128-byte swizzle, OOB reads zero-filled / OOB writes clipped.
int make_tmap_2d(CUtensorMap* m, const void* base, bool is_f32, long long rows, long long cols, long long ld,
                        int box_rows) {
  if (resolve_encode()) return fail("cuTensorMapEncodeTiled entry point not available");
  const int esz = is_f32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_encode(m, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed, CUresult=" + std::to_string((int)r));
  return 0;
}

static int g_gemm_ctas = 0;  // 0 = auto, 1 = single-CTA tiles, 2 = CTA-pair (cta_group::2) tiles
void set_gemm_ctas(int v) { g_gemm_ctas = v; }

// 4-D NHWC bf16 activation tensor [B,H,W,C] for the implicit-GEMM conv: box = [64 ch, w_box, h_box, 1], 128B swizzle
static int make_tmap_nhwc(CUtensorMap* m, const void* base, int B, int H, int W, int C, int w_box, int h_box) {
  if (resolve_encode()) return fail("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)w_box, (cuuint32_t)h_box, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(4d) failed, CUresult=" + std::to_string((int)r));
  return 0;
}

// 3-D view [B][positions][C] of a token-major bf16 activation: box = [64 ch, box_pos, 1], 128B swizzle.  Positions outside
// [0, npos) of the SAME image are zero-filled (a 2-D [B*npos, C] view would read the neighbouring image instead).
int make_tmap_tokens3d(CUtensorMap* m, const void* base, int B, int npos, int C, int box_pos) {
  if (resolve_encode()) return fail("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)npos, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)npos * C * 2};
  cuuint32_t box[3] = {64u, (cuuint32_t)box_pos, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(3d) failed, CUresult=" + std::to_string((int)r));
  return 0;
}

template <int BN, int EPI, int CTAS>
static int launch_t(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& td, int M, int N,
                    int K, const GemmEpi& ep, cudaStream_t st) {
  auto kern = gemm_bf16_tn_kernel<BN, EPI, CTAS>;
  constexpr int smem = GemmSmem<BN, EPI, CTAS>::TOTAL;
  static bool attr_set = false;
  if (!attr_set) {
    TLD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int tiles = ((M + GEMM_BM * CTAS - 1) / (GEMM_BM * CTAS)) * ((N + BN - 1) / BN);
  const int slots = sm_count() / CTAS;  // CTAs (or CTA pairs) resident at once
  const int grid = (tiles < slots ? tiles : slots) * CTAS;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  TLD_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, td, M, N, K, ep));
  return 0;
}

template <int BN, int CTAS>
static int launch_bn(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& td, int M,
                     int N, int K, const GemmEpi& ep, cudaStream_t st) {
  switch (epi) {
    case EPI_BF16: return launch_t<BN, EPI_BF16, CTAS>(ta, tb, tc, td, M, N, K, ep, st);
    case EPI_BIAS_BF16: return launch_t<BN, EPI_BIAS_BF16, CTAS>(ta, tb, tc, td, M, N, K, ep, st);
    case EPI_BIAS_RESID_F32: return launch_t<BN, EPI_BIAS_RESID_F32, CTAS>(ta, tb, tc, td, M, N, K, ep, st);
    case EPI_XATTN_RESID_F32: return launch_t<BN, EPI_XATTN_RESID_F32, CTAS>(ta, tb, tc, td, M, N, K, ep, st);
    case EPI_F32: return launch_t<BN, EPI_F32, CTAS>(ta, tb, tc, td, M, N, K, ep, st);
    case EPI_LNFOLD_BF16: return launch_t<BN, EPI_LNFOLD_BF16, CTAS>(ta, tb, tc, td, M, N, K, ep, st);
    case EPI_BIAS_RESID_LNP: return launch_t<BN, EPI_BIAS_RESID_LNP, CTAS>(ta, tb, tc, td, M, N, K, ep, st);
    case EPI_XATTN_RESID_LNP: return launch_t<BN, EPI_XATTN_RESID_LNP, CTAS>(ta, tb, tc, td, M, N, K, ep, st);
  }
  return fail("launch_gemm: unknown epilogue " + std::to_string(epi));
}

// Tile width: the widest BN in {256,192,128,64} that divides N, preferring the one with the least
// wave-quantisation loss on the persistent grid (ties -> wider tile, fewer A re-reads).
static int pick_bn(int M, int N, int ctas) {
  const int cands[4] = {256, 192, 128, 64};
  const int sms = sm_count() / ctas;
  const int m_tiles = (M + GEMM_BM * ctas - 1) / (GEMM_BM * ctas);
  int best = 0;
  double best_cost = 1e30;
  for (int bn : cands) {
    if (N % bn != 0 && !(bn == 64)) continue;
    const int n_tiles = (N + bn - 1) / bn;
    const long long tiles = (long long)m_tiles * n_tiles;
    const long long waves = (tiles + sms - 1) / sms;
    // time ~ waves * (tile work ~ bn) / efficiency: a 64-wide tile re-reads its A tile four times as often as a 256-wide one
    // and runs at roughly 0.6 of the wide tiles' MMA rate (measured on the VAE's 512-channel convolutions: 744 TFLOP/s with
    // BN = 64 against 1.5 PFLOP/s with BN = 256), a 128-wide one at ~0.9
    const double eff = bn >= 192 ? 1.0 : (bn == 128 ? 0.92 : 0.62);
    const double cost = double(waves) * bn * (1.0 + 8.0 / bn) / eff;
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

static int dispatch(int ctas, int bn, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                    const CUtensorMap& td, int M, int N, int K, const GemmEpi& ep, cudaStream_t st);

// 3x3 'same' convolution as an implicit GEMM on the tcgen05 core: x NHWC bf16 [B,H,W,Cin], w bf16 [Cout, 9*Cin]
// (K order = (ky, kx, cin)), bias fp32 [Cout] -> out NHWC bf16 [B,H,W,Cout].
int launch_conv3x3(const bf16* x, const bf16* w, const float* bias, bf16* out, int B, int H, int W, int Cin, int Cout,
                   cudaStream_t st, const bf16* resid, float* gn_part) {
  TLD_CHECK(Cin % 64 == 0 && Cout % 64 == 0, "conv3x3: Cin and Cout must be multiples of 64");
  TLD_CHECK((H * W) % 128 == 0, "conv3x3: H*W must be a multiple of 128");
  const int w_box = W < 128 ? W : 128;
  TLD_CHECK(128 % w_box == 0 && W % w_box == 0, "conv3x3: width must divide 128 or be a multiple of 128");
  const int h_box = 128 / w_box;
  TLD_CHECK(H % h_box == 0, "conv3x3: height must be a multiple of 128/width");
  const long long M = (long long)B * H * W;
  TLD_CHECK(M < (1LL << 31), "conv3x3: too many pixels");
  const int K = 9 * Cin, N = Cout;
  const int ctas = g_gemm_ctas ? g_gemm_ctas : (M >= 4096 ? 2 : 1);
  const int bn = pick_bn((int)M, N, ctas);
  CUtensorMap ta, tb, tc;
  if (make_tmap_nhwc(&ta, x, B, H, W, Cin, w_box, h_box)) return 1;
  if (make_tmap_2d(&tb, w, false, N, K, K, bn / ctas)) return 1;
  if (make_tmap_2d(&tc, out, false, M, N, N, 32)) return 1;
  TLD_CHECK((resid == nullptr && gn_part == nullptr) || bias != nullptr, "conv3x3: the residual / GroupNorm-partials epilogue needs a bias");
  TLD_CHECK((reinterpret_cast<uintptr_t>(resid) & 15) == 0 && (reinterpret_cast<uintptr_t>(gn_part) & 7) == 0,
            "conv3x3: residual must be 16-byte aligned");
  GemmEpi ep{};
  ep.bias = bias;
  ep.conv_cpb = Cin / 64;
  ep.conv_h = H;
  ep.conv_w = W;
  ep.resid = resid;
  ep.gn_part = gn_part;
  return dispatch(ctas, bn, bias ? EPI_BIAS_BF16 : EPI_BF16, ta, tb, tc, tc, (int)M, N, K, ep, st);
}

int launch_gemm(int epi, const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, void* out, int ldo,
                const float* bias, const XattnArgs* xa, cudaStream_t st, const LnFoldArgs* ln) {
  TLD_CHECK(M > 0 && N > 0 && K > 0, "launch_gemm: empty problem");
  TLD_CHECK(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "launch_gemm: K/lda/ldw must be multiples of 8 (16-byte TMA rows)");
  TLD_CHECK(N % 32 == 0, "launch_gemm: N must be a multiple of 32");
  TLD_CHECK((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
            "launch_gemm: operands must be 16-byte aligned");
  if (epi == EPI_XATTN_RESID_F32 || epi == EPI_XATTN_RESID_LNP) {
    TLD_CHECK(xa != nullptr, "launch_gemm: cross-attention epilogue needs XattnArgs");
    TLD_CHECK(N % 64 == 0 && xa->n_tok % 64 == 0, "launch_gemm: cross-attention epilogue needs N and n_tok multiples of 64");
    TLD_CHECK(xa->kv0_stride % 4 == 0 && xa->kv1_stride % 4 == 0 && xa->embed_dim % 4 == 0 &&
                  (reinterpret_cast<uintptr_t>(xa->kv0) & 15) == 0 && (reinterpret_cast<uintptr_t>(xa->kv1) & 15) == 0,
              "launch_gemm: cross-attention K/V rows must be 16-byte aligned (float4 staging)");
  }
  if (epi == EPI_BIAS_BF16 || epi == EPI_BIAS_RESID_F32 || epi == EPI_LNFOLD_BF16 || epi == EPI_BIAS_RESID_LNP)
    TLD_CHECK(bias != nullptr, "launch_gemm: bias epilogues need a bias");
  const bool lnp = epi == EPI_BIAS_RESID_LNP || epi == EPI_XATTN_RESID_LNP;
  if (epi == EPI_LNFOLD_BF16)
    TLD_CHECK(ln && ln->col_s && ln->row_part && ln->n_part == K / 32 && K % 64 == 0 && K <= 1024 &&
                  (reinterpret_cast<uintptr_t>(ln->row_part) & 15) == 0,
              "launch_gemm: the LayerNorm-fold epilogue needs col_s, 16-byte aligned row partials [M, K/32] and K % 64 == 0");
  if (lnp)
    TLD_CHECK(ln && ln->xb_out && ln->part_out && N % 64 == 0 && ldo == N && ln->ldxb % 8 == 0 &&
                  (reinterpret_cast<uintptr_t>(ln->xb_out) & 15) == 0,
              "launch_gemm: the producer epilogues need xb_out, part_out, a dense fp32 output (ldo == N) and N % 64 == 0");
  const bool out_f32 = !(epi == EPI_BF16 || epi == EPI_BIAS_BF16 || epi == EPI_LNFOLD_BF16);
  TLD_CHECK((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (ldo * (out_f32 ? 4 : 2)) % 16 == 0,
            "launch_gemm: output must be 16-byte aligned with a 16-byte multiple row pitch");
  // CTA pairs pay off once there are enough 256-row tiles to fill the 74 SM pairs
  const int ctas = g_gemm_ctas ? g_gemm_ctas : (M >= 4096 ? 2 : 1);
  int bn = pick_bn(M, N, ctas);
  if (lnp && ctas == 1 && bn > 128) bn = (N % 128 == 0) ? 128 : 64;   // 96 KB of epilogue slabs: keep >= 3 pipeline stages
  CUtensorMap ta, tb, tc, td;
  if (make_tmap_2d(&ta, A, false, M, K, lda, GEMM_BM)) return 1;
  if (make_tmap_2d(&tb, W, false, N, K, ldw, bn / ctas)) return 1;
  if (make_tmap_2d(&tc, out, out_f32, M, N, ldo, 32)) return 1;
  td = tc;
  if (lnp && make_tmap_2d(&td, ln->xb_out, false, M, N, ln->ldxb, 32)) return 1;
  GemmEpi ep{};
  ep.bias = bias;
  ep.scale = 0.125f;  // 1/sqrt(64)
  if (ln) {
    ep.col_s = ln->col_s;
    ep.row_part = ln->row_part;
    ep.n_part = ln->n_part;
    ep.inv_d = 1.f / float(K);
    ep.ln_eps = ln->ln_eps;
    ep.x_in = reinterpret_cast<const float*>(out);
    ep.part_out = ln->part_out;
  }
  if (xa) {
    ep.kv0 = xa->kv0;
    ep.kv1 = xa->kv1;
    ep.kv0_stride = xa->kv0_stride;
    ep.kv1_stride = xa->kv1_stride;
    ep.step_ptr = xa->step_ptr;
    ep.n_tok = xa->n_tok;
    ep.embed_dim = xa->embed_dim;
  }
  return dispatch(ctas, bn, epi, ta, tb, tc, td, M, N, K, ep, st);
}

// MN-major operand modes (see GemmEpi::mn_major).  a_mn: A stored [K, M] instead of [M, K]; b_mn: B stored [K, N] instead of
// [N, K]; lda / ldb = row pitch of the stored matrix.  out fp32 (EPI_F32) or bf16 (EPI_BF16).
static int launch_gemm_major(int epi, bool a_mn, bool b_mn, const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K,
                             void* out, int ldo, cudaStream_t st) {
  TLD_CHECK(epi == EPI_F32 || epi == EPI_BF16, "launch_gemm_mn/nn: only the plain fp32 / bf16 epilogues");
  TLD_CHECK(M > 0 && N > 0 && K > 0, "launch_gemm_mn/nn: empty problem");
  TLD_CHECK(M % 8 == 0 && K % 8 == 0 && N % 64 == 0 && lda % 8 == 0 && ldb % 8 == 0,
            "launch_gemm_mn/nn: M/K/lda/ldb multiples of 8 and N a multiple of 64");
  TLD_CHECK((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
            "launch_gemm_mn/nn: operands must be 16-byte aligned");
  const bool out_f32 = epi == EPI_F32;
  TLD_CHECK((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (ldo * (out_f32 ? 4 : 2)) % 16 == 0,
            "launch_gemm_mn/nn: output must be 16-byte aligned with a 16-byte multiple row pitch");
  // a CTA pair splits the N tile in two: with an MN-major B each half must be whole 64-column atoms
  int ctas = g_gemm_ctas ? g_gemm_ctas : (M >= 4096 ? 2 : 1);
  int bn = pick_bn(M, N, ctas);
  if (b_mn && ctas == 2 && (bn / 2) % 64 != 0) {
    if (N % 256 == 0 || N % 128 == 0) {
      bn = N % 256 == 0 ? 256 : 128;   // keep the pair, take a tile whose halves are whole atoms
    } else {
      ctas = 1;
      bn = pick_bn(M, N, 1);
    }
  }
  CUtensorMap ta, tb, tc;
  if (a_mn ? make_tmap_2d(&ta, A, false, K, M, lda, 64) : make_tmap_2d(&ta, A, false, M, K, lda, GEMM_BM)) return 1;
  if (b_mn ? make_tmap_2d(&tb, B, false, K, N, ldb, 64) : make_tmap_2d(&tb, B, false, N, K, ldb, bn / ctas)) return 1;
  if (make_tmap_2d(&tc, out, out_f32, M, N, ldo, 32)) return 1;
  GemmEpi ep{};
  ep.mn_major = (a_mn ? 1 : 0) | (b_mn ? 2 : 0);
  return dispatch(ctas, bn, epi, ta, tb, tc, tc, M, N, K, ep, st);
}
// C[M,N] = A^T B: the weight-gradient shape dW = dY^T X (K = tokens)
int launch_gemm_mn(int epi, const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K, void* out, int ldo,
                   cudaStream_t st) {
  return launch_gemm_major(epi, true, true, A, lda, B, ldb, M, N, K, out, ldo, st);
}
// C[M,N] = A B with B stored [K, N]: the data-gradient shape dX = dY W with the weight as it is
int launch_gemm_nn(int epi, const bf16* A, int lda, const bf16* B, int ldb, int M, int N, int K, void* out, int ldo,
                   cudaStream_t st) {
  return launch_gemm_major(epi, false, true, A, lda, B, ldb, M, N, K, out, ldo, st);
}

static int dispatch(int ctas, int bn, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                    const CUtensorMap& td, int M, int N, int K, const GemmEpi& ep, cudaStream_t st) {
  if (ctas == 2) {
    switch (bn) {
      case 256: return launch_bn<256, 2>(epi, ta, tb, tc, td, M, N, K, ep, st);
      case 192: return launch_bn<192, 2>(epi, ta, tb, tc, td, M, N, K, ep, st);
      case 128: return launch_bn<128, 2>(epi, ta, tb, tc, td, M, N, K, ep, st);
      case 64: return launch_bn<64, 2>(epi, ta, tb, tc, td, M, N, K, ep, st);
    }
  }
  switch (bn) {
    case 256: return launch_bn<256, 1>(epi, ta, tb, tc, td, M, N, K, ep, st);
    case 192: return launch_bn<192, 1>(epi, ta, tb, tc, td, M, N, K, ep, st);
    case 128: return launch_bn<128, 1>(epi, ta, tb, tc, td, M, N, K, ep, st);
    case 64: return launch_bn<64, 1>(epi, ta, tb, tc, td, M, N, K, ep, st);
  }
  return fail("gemm dispatch: no tile width for N=" + std::to_string(N));
}

}  // namespace tld
