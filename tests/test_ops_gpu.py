"""Single-kernel parity: every CUDA kernel of libtld_b200 (through its C-ABI `tld_op_*` entry point) against a
plain PyTorch fp32 statement of the same op on the same inputs.

Tolerances: operands are bf16-rounded BEFORE the reference is computed, accumulation is fp32 on both sides, so
fp32-output ops must agree to ~1e-5 relative; bf16-output ops to one bf16 ulp (2^-8 relative).
"""
import math

import pytest
import torch

from conftest import rel_fro

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from transformer_latent_diffusion_b200 import _lib

    return _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _bf16_bits(t):
    return t.view(torch.int16)


def _err_map(got, ref, bm=32, bn=32):
    """coarse map of where the error lives (helps diagnosing tile/descriptor bugs)"""
    d = (got.double() - ref.double()).abs()
    M, N = d.shape
    rows = []
    for i in range(0, min(M, 256), bm):
        rows.append(" ".join(f"{d[i:i + bm, j:j + bn].max().item():8.2e}" for j in range(0, min(N, 256), bn)))
    return "\n".join(rows)


GEMM_SHAPES = [
    # (M, N, K)
    (128, 64, 64), (128, 128, 64), (128, 256, 128), (256, 192, 256), (128, 768, 768),
    (384, 2304, 768), (512, 3072, 768), (256, 768, 3072), (200, 256, 256), (64, 384, 128),
    (4096, 768, 768),
]


@pytest.fixture(params=[1, 2], ids=["cta1", "cta2"])
def ctas(request, lib):
    """run the GEMM tests on single-CTA tiles and on CTA-pair (cta_group::2) tiles"""
    lib.check(lib.load().tld_set_option(b"gemm_ctas", request.param), "set_option")
    yield request.param
    lib.check(lib.load().tld_set_option(b"gemm_ctas", 0), "set_option")


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("epi", [0, 4])
def test_gemm_plain(lib, ctas, M, N, K, epi):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    ref = A.float() @ W.float().t()
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16 if epi == 0 else torch.float32)
    lib.check(lib.load().tld_op_gemm(epi, lib.ptr(A), lib.ptr(W), M, N, K, lib.ptr(out), None, _stream()), "gemm")
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all(), "non-finite / unwritten outputs\n" + _err_map(out.float().nan_to_num(1e9), ref)
    err = rel_fro(out.float(), ref)
    tol = 4e-3 if epi == 0 else 2e-5
    assert err < tol, f"rel_fro={err:.3e}\n" + _err_map(out.float(), ref)


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (768, 3072, 1024), (3072, 768, 8192), (2304, 768, 520), (200, 128, 136),
                                   (4096, 1024, 256)])
def test_gemm_mn_major(lib, ctas, M, N, K):
    """dW = dY^T X on MN-major operands (A [K,M], B [K,N] row-major): the wgrad shape, incl. ragged M and K."""
    g = torch.Generator(device="cuda").manual_seed(11)
    A = (torch.randn(K, M, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    B = torch.randn(K, N, device="cuda", generator=g).bfloat16()
    ref = A.float().t() @ B.float()
    out = torch.full((M, N), float("nan"), device="cuda")
    lib.check(lib.load().tld_op_gemm_mn(4, lib.ptr(A), lib.ptr(B), M, N, K, lib.ptr(out), _stream()), "gemm_mn")
    torch.cuda.synchronize()
    assert torch.isfinite(out).all(), "non-finite / unwritten outputs\n" + _err_map(out.nan_to_num(1e9), ref)
    assert rel_fro(out, ref) < 2e-5, _err_map(out, ref)


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (8192, 3072, 768), (4096, 768, 3072), (1000, 768, 2304), (256, 192, 72)])
@pytest.mark.parametrize("epi", [0, 4])
def test_gemm_nn_mn_major_b(lib, ctas, M, N, K, epi):
    """dX = dY W with W stored [K, N] (the weight as it is): K-major A, MN-major B - the dgrad shape, no transposed weight"""
    g = torch.Generator(device="cuda").manual_seed(12)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(K, N, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    ref = A.float() @ B.float()
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16 if epi == 0 else torch.float32)
    lib.check(lib.load().tld_op_gemm_nn(epi, lib.ptr(A), lib.ptr(B), M, N, K, lib.ptr(out), _stream()), "gemm_nn")
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all(), "non-finite / unwritten outputs"
    assert rel_fro(out.float(), ref) < (4e-3 if epi == 0 else 2e-5), _err_map(out.float(), ref)


@pytest.mark.parametrize("M,N,K", [(256, 1024, 256), (384, 3072, 768), (256, 160, 128), (300, 96, 64), (128, 32, 64)])
def test_gemm_bias_bf16(lib, ctas, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    ref = A.float() @ W.float().t() + bias
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    lib.check(lib.load().tld_op_gemm(1, lib.ptr(A), lib.ptr(W), M, N, K, lib.ptr(out), lib.ptr(bias), _stream()), "gemm")
    assert rel_fro(out.float(), ref) < 4e-3


@pytest.mark.parametrize("M,N,K", [(256, 256, 1024), (384, 768, 3072), (130, 128, 512), (256, 160, 128), (140, 96, 64)])
def test_gemm_bias_residual(lib, ctas, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(2)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g)
    ref = x + A.float() @ W.float().t() + bias
    lib.check(lib.load().tld_op_gemm(2, lib.ptr(A), lib.ptr(W), M, N, K, lib.ptr(x), lib.ptr(bias), _stream()), "gemm")
    assert rel_fro(x, ref) < 2e-5, _err_map(x, ref)


@pytest.mark.parametrize("B,n_tok,D", [(2, 64, 128), (3, 64, 128), (2, 256, 256), (3, 256, 768), (1, 1024, 128)])
def test_gemm_cross_attention_epilogue(lib, ctas, B, n_tok, D):
    """q_linear + 2-key SDPA + residual (transformer_blocks.py:70-72,137) fused in the GEMM epilogue."""
    g = torch.Generator(device="cuda").manual_seed(3)
    M = B * n_tok
    A = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    Wq = (torch.randn(D, D, device="cuda", generator=g) / math.sqrt(D)).bfloat16()
    kv0 = torch.randn(B, 2 * D, device="cuda", generator=g)
    kv1 = torch.randn(B, 2 * D, device="cuda", generator=g)
    x = torch.randn(M, D, device="cuda", generator=g)
    H = D // 64
    q = (A.float() @ Wq.float().t()).view(B, n_tok, H, 64)
    k = torch.stack([kv0[:, :D], kv1[:, :D]], 1).view(B, 2, H, 64)
    v = torch.stack([kv0[:, D:], kv1[:, D:]], 1).view(B, 2, H, 64)
    s = torch.einsum("bnhd,bshd->bhns", q, k) / 8.0
    o = torch.einsum("bhns,bshd->bnhd", torch.softmax(s, -1), v).reshape(M, D)
    ref = x + o
    lib.check(lib.load().tld_op_gemm_xattn(lib.ptr(A), lib.ptr(Wq), M, D, lib.ptr(x), lib.ptr(kv0), lib.ptr(kv1), n_tok,
                                           _stream()), "gemm_xattn")
    assert rel_fro(x, ref) < 1e-4, _err_map(x, ref)


@pytest.mark.parametrize("rows,D", [(8, 128), (13, 256), (1000, 768), (5, 1024)])
def test_layernorm(lib, rows, D):
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(rows, D, device="cuda", generator=g) * 3 + 1.5
    w = torch.randn(D, device="cuda", generator=g)
    b = torch.randn(D, device="cuda", generator=g)
    ref = torch.nn.functional.layer_norm(x, (D,), w, b, 1e-5)
    y = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    lib.check(lib.load().tld_op_layernorm(lib.ptr(x), lib.ptr(w), lib.ptr(b), lib.ptr(y), rows, D, _stream()), "ln")
    assert (y.float() - ref).abs().max() <= 2 ** -8 * ref.abs().max() + 1e-6
    assert rel_fro(y.float(), ref) < 3e-3


@pytest.mark.parametrize("impl", [1, 3])
@pytest.mark.parametrize("B,n_tok,D", [(1, 64, 128), (2, 256, 128), (3, 256, 768), (1, 1024, 256), (1, 4096, 128),
                                       (40, 128, 768), (17, 256, 768), (2, 1024, 768),   # > 2 tiles per persistent CTA
                                       (3, 384, 192), (50, 384, 320)])                   # 3 q-tiles / 3 and 5 heads: multiply-high tile split
def test_self_attention(lib, B, n_tok, D, impl):
    if impl >= 2 and n_tok % 128:
        pytest.skip("tcgen05 attention needs n_tok % 128 == 0")
    g = torch.Generator(device="cuda").manual_seed(5)
    T = B * n_tok
    qkv = torch.randn(T, 3 * D, device="cuda", generator=g).bfloat16()
    x = torch.randn(T, D, device="cuda", generator=g)
    H = D // 64
    q, k, v = (t.float().view(B, n_tok, H, 64).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=1))
    s = (q @ k.transpose(-1, -2)) / 8.0
    o = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(T, D)
    ref = x + o
    lib.check(lib.load().tld_op_self_attention(lib.ptr(qkv), lib.ptr(x), B, n_tok, D, impl, _stream()), "attn")
    # P is rounded to bf16 before the PV product (as in every flash kernel): error ~2^-9 relative on o
    assert rel_fro(x - (ref - o), o) < 6e-3, _err_map(x, ref)


# Variants of the norm2 + cross-attention + norm3 row kernel: (xattn_mma, xattn_rows, xattn_ctas).
#   xattn_mma 0 = FFMA row kernel, 1 / 2 / 3 = dot products as tf32 mma.sync with x truncated / x split / x and keys split
#   xattn_rows = rows per warp of the FFMA kernel (4 or 2), xattn_ctas = its CTAs per SM (1 or 2)
# Tolerance on the attention output (relative Frobenius, fp64 reference): the logits see the folded keys rounded to tf32 in
# mma modes 1 / 2 (2^-12 per element) and x truncated to tf32 in mode 1 (2^-10); the FFMA kernels and mode 3 are fp32-level.
_XL_VARIANTS = {"ffma_r4c2": (0, 4, 2, 2e-5), "ffma_r4c1": (0, 4, 1, 2e-5), "ffma_r2c1": (0, 2, 1, 2e-5), "ffma_r2c2": (0, 2, 2, 2e-5),
                "mma_trunc": (1, 4, 1, 2e-3), "mma_xsplit": (2, 4, 1, 3e-4), "mma_split": (3, 4, 1, 2e-5)}
_XL_DEFAULT = (0, 4, 1)
_XL_SHAPES = [(2, 64, 128), (3, 256, 768), (5, 32, 256), (37, 256, 768), (2, 1024, 384), (4, 96, 1024), (3, 36, 512), (9, 256, 896),
              (148, 256, 768), (7, 40, 640), (3, 8, 512), (120, 64, 256), (600, 32, 128), (30, 1024, 256)]
_XL_FEW = [(3, 256, 768), (5, 32, 256), (148, 256, 768), (4, 96, 1024), (7, 40, 640)]


@pytest.mark.parametrize("variant,B,n_tok,D", [(v, *sh) for v in ("ffma_r4c1", "mma_xsplit") for sh in _XL_SHAPES] +
                         [(v, *sh) for v in ("ffma_r4c2", "ffma_r2c1", "ffma_r2c2", "mma_trunc", "mma_split") for sh in _XL_FEW])
def test_ln_xattn_ln_fused(lib, B, n_tok, D, variant):
    """norm2 + 2-token cross-attention (q folded into the keys) + residual + norm3 in one row-wise kernel vs fp32 torch.
    Shapes the tensor-pipe variant does not take (embed_dim > 768, tokens % 8 != 0) run the FFMA kernel in every mode.
    The shapes cover CTAs whose row range lies inside one sample, spans several samples (a table build per sample) and
    grids smaller than the GPU."""
    mma, rows, ctas, tol = _XL_VARIANTS[variant]
    L = lib.load()
    for k, v in ((b"xattn_mma", mma), (b"xattn_rows", rows), (b"xattn_ctas", ctas)):
        lib.check(L.tld_set_option(k, v), "opt")
    try:
        _ln_xattn_ln_case(lib, B, n_tok, D, tol)
    finally:
        for k, v in zip((b"xattn_mma", b"xattn_rows", b"xattn_ctas"), _XL_DEFAULT):
            lib.check(L.tld_set_option(k, v), "opt")


def _ln_xattn_ln_case(lib, B, n_tok, D, tol):
    g = torch.Generator(device="cuda").manual_seed(B * 100 + D)
    T, H = B * n_tok, D // 64
    x = torch.randn(T, D, device="cuda", generator=g) * 2 + 0.3
    g2, b2, g3, b3 = (torch.randn(D, device="cuda", generator=g) * s + o for s, o in ((0.2, 1.0), (0.2, 0.0), (0.2, 1.0), (0.2, 0.0)))
    wq = (torch.randn(D, D, device="cuda", generator=g) / D ** 0.5).bfloat16()
    kv0 = torch.randn(B, 2 * D, device="cuda", generator=g)
    kv1 = torch.randn(B, 2 * D, device="cuda", generator=g)
    xn2 = torch.nn.functional.layer_norm(x.double(), (D,), g2.double(), b2.double())
    q = (xn2 @ wq.double().t()).view(B, n_tok, H, 64)
    k = torch.stack([kv0[:, :D], kv1[:, :D]], 1).double().view(B, 2, H, 64)
    v = torch.stack([kv0[:, D:], kv1[:, D:]], 1).double().view(B, 2, H, 64)
    p = torch.softmax(torch.einsum("bnhd,bjhd->bnhj", q, k) / 8.0, -1)
    o = torch.einsum("bnhj,bjhd->bnhd", p, v).reshape(T, D)
    x_ref = x.double() + o
    y_ref = torch.nn.functional.layer_norm(x_ref, (D,), g3.double(), b3.double())
    x0 = x.clone()
    uk = torch.empty(2 * B, H, D, device="cuda")
    y = torch.empty(T, D, device="cuda", dtype=torch.bfloat16)
    lib.check(lib.load().tld_op_ln_xattn_ln(lib.ptr(x), lib.ptr(g2), lib.ptr(b2), lib.ptr(g3), lib.ptr(b3), lib.ptr(wq), lib.ptr(kv0),
                                            lib.ptr(kv1), B, n_tok, D, lib.ptr(uk), lib.ptr(y), _stream()), "ln_xattn_ln")
    torch.cuda.synchronize()
    assert torch.isfinite(x).all() and torch.isfinite(y.float()).all()
    assert rel_fro((x - x0).double(), o) < tol, _err_map(x, x_ref.float())
    assert rel_fro(y.double(), y_ref) < 4e-3


@pytest.mark.parametrize("B,D", [(1, 64), (2, 128), (3, 768), (5, 192), (40, 768), (75, 320), (128, 768)])
def test_qkv_attention_fused(lib, B, D):
    """qkv projection + attention + residual add in one CTA-pair kernel (256 tokens per sample) vs fp32 math on bf16-rounded q, k, v"""
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + D)
    n_tok, T, H = 256, B * 256, D // 64
    xn = torch.randn(T, D, device="cuda", generator=g).bfloat16()
    w = (torch.randn(3 * D, D, device="cuda", generator=g) * (1.5 / D ** 0.5)).bfloat16()
    x = torch.randn(T, D, device="cuda", generator=g)
    qkv = (xn.float() @ w.float().t()).bfloat16()
    q, k, v = (t.float().view(B, n_tok, H, 64).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=1))
    o = (torch.softmax((q @ k.transpose(-1, -2)) / 8.0, -1) @ v).permute(0, 2, 1, 3).reshape(T, D)
    x0 = x.clone()
    lib.check(lib.load().tld_op_qkv_attention(lib.ptr(xn), lib.ptr(w), lib.ptr(x), B, n_tok, D, _stream()), "qkv_attention")
    torch.cuda.synchronize()
    assert torch.isfinite(x).all()
    assert rel_fro(x - x0, o) < 6e-3, _err_map(x, x0 + o)
    # and against the two-kernel path (GEMM -> qkv in HBM -> attention_tc2) on the same operands
    x2 = x0.clone()
    lib.check(lib.load().tld_op_self_attention(lib.ptr(qkv), lib.ptr(x2), B, n_tok, D, 3, _stream()), "attn")
    assert rel_fro(x - x0, x2 - x0) < 6e-3


@pytest.mark.parametrize("emu", [0, 4, 6, 8, 10])
@pytest.mark.parametrize("qk_scale", [1.0, 4.0])
def test_self_attention_persistent_variants(lib, emu, qk_scale):
    """attention_tc2: every FMA-pipe exp2 share, and (qk_scale 4: score std ~23 in the exp2 domain) the lazy-rescale path
    where the row maximum keeps jumping by more than 2^8 between 64-key chunks."""
    B, n_tok, D = 5, 512, 256
    g = torch.Generator(device="cuda").manual_seed(50 + emu)
    T = B * n_tok
    qkv = torch.randn(T, 3 * D, device="cuda", generator=g)
    qkv[:, : 2 * D] *= qk_scale
    qkv = qkv.bfloat16()
    x = torch.zeros(T, D, device="cuda")
    H = D // 64
    q, k, v = (t.float().view(B, n_tok, H, 64).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=1))
    ref = (torch.softmax((q @ k.transpose(-1, -2)) / 8.0, -1) @ v).permute(0, 2, 1, 3).reshape(T, D)
    L = lib.load()
    lib.check(L.tld_set_option(b"attention_exp_emu", emu), "opt")
    try:
        lib.check(L.tld_op_self_attention(lib.ptr(qkv), lib.ptr(x), B, n_tok, D, 3, _stream()), "attn")
        torch.cuda.synchronize()
    finally:
        lib.check(L.tld_set_option(b"attention_exp_emu", 6), "opt")
    assert rel_fro(x, ref) < 6e-3, _err_map(x, ref)


@pytest.mark.parametrize("B,grid,C", [(1, 8, 512), (2, 16, 1024), (2, 16, 3072), (1, 32, 512), (3, 32, 128), (2, 64, 256),
                                      (1, 64, 3072), (2, 32, 36)])
def test_dwconv_gelu(lib, B, grid, C):
    g = torch.Generator(device="cuda").manual_seed(6)
    h = torch.randn(B, grid, grid, C, device="cuda", generator=g).bfloat16()
    w = torch.randn(C, 1, 3, 3, device="cuda", generator=g) / 3
    bias = torch.randn(C, device="cuda", generator=g) * 0.1
    ref = torch.nn.functional.conv2d(h.float().permute(0, 3, 1, 2), w, bias, padding=1, groups=C)
    ref = torch.nn.functional.gelu(ref).permute(0, 2, 3, 1).contiguous()
    w9 = w.view(C, 9).t().contiguous()
    out = torch.empty_like(h)
    lib.check(lib.load().tld_op_dwconv_gelu(lib.ptr(h), lib.ptr(w9), lib.ptr(bias), lib.ptr(out), B, grid, C, _stream()),
              "dwconv")
    assert rel_fro(out.float(), ref) < 3e-3
    assert (out.float() - ref).abs().max() <= 2 ** -8 * ref.abs().max() + 1e-5


@pytest.mark.parametrize("fold", [False, True])
@pytest.mark.parametrize("B,K,N", [(1, 64, 256), (3, 768, 3072), (80, 128, 512), (37, 256, 1024)])
def test_gemm_up_dwconv_gelu_fused(lib, B, K, N, fold):
    """MLPSepConv front half in one kernel (transformer_blocks.py:95-103): up-projection (tcgen05 CTA-pair tile = one 16x16
    image) + depthwise 3x3 + bias + GELU with the hidden tensor kept on chip; halo row exchanged between the two CTAs of a
    pair through distributed shared memory.  80 / 37 images: several tiles per persistent CTA pair (barrier phases wrap).
    fold: LayerNorm folded into the GEMM (gamma into the weight, mean / rstd applied on the accumulator)."""
    g = torch.Generator(device="cuda").manual_seed(B * 31 + K)
    M = B * 256
    W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    bias = torch.randn(N, device="cuda", generator=g) * 0.2
    dw = torch.randn(N, 1, 3, 3, device="cuda", generator=g) / 3
    dwb = torch.randn(N, device="cuda", generator=g) * 0.1
    w9 = dw.view(N, 9).t().contiguous()
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    if fold:
        x = torch.randn(M, K, device="cuda", generator=g) * 1.7 + 0.3 * torch.randn(M, 1, device="cuda", generator=g)
        gamma = 1 + 0.2 * torch.randn(K, device="cuda", generator=g)
        beta = 0.1 * torch.randn(K, device="cuda", generator=g)
        hid = torch.nn.functional.layer_norm(x, (K,), gamma, beta, 1e-5) @ W.t() + bias            # fp32 statement
        A = x.bfloat16()
        Wf = (W * gamma).bfloat16()
        col_s = Wf.float().sum(1).contiguous()
        col_c = (W @ beta + bias).contiguous()
        sums = torch.stack([x.sum(1), (x * x).sum(1)], 1).contiguous()
        args = (lib.ptr(A), lib.ptr(Wf), lib.ptr(col_c), lib.ptr(col_s), lib.ptr(sums))
    else:
        A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        Wf = W.bfloat16()
        hid = A.float() @ Wf.float().t() + bias
        args = (lib.ptr(A), lib.ptr(Wf), lib.ptr(bias), None, None)
    hid = hid.bfloat16().float().view(B, 16, 16, N)              # the kernel rounds the hidden tensor to bf16 like the 3-kernel path
    ref = torch.nn.functional.conv2d(hid.permute(0, 3, 1, 2), dw, dwb, padding=1, groups=N)
    ref = torch.nn.functional.gelu(ref).permute(0, 2, 3, 1).reshape(M, N)
    lib.check(lib.load().tld_op_gemm_up_dwconv_gelu(*args, lib.ptr(w9), lib.ptr(dwb), lib.ptr(out), B, K, N, _stream()),
              "gemm_up_dwconv_gelu")
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all(), "non-finite / unwritten outputs\n" + _err_map(out.float().nan_to_num(1e9), ref)
    err = rel_fro(out.float(), ref)
    assert err < (1.2e-2 if fold else 6e-3), f"rel_fro={err:.3e}\n" + _err_map(out.float(), ref)


def _row_partials(x):
    M, D = x.shape
    g = x.view(M, D // 32, 32)
    return torch.stack([g.sum(-1), (g * g).sum(-1)], -1).contiguous()     # [M, D/32, 2]


@pytest.mark.parametrize("M,N,K", [(256, 2304, 768), (4096, 768, 256), (300, 192, 128), (8192, 3072, 768)])
def test_gemm_layernorm_fold_consumer(lib, ctas, M, N, K):
    """norm1 / norm3 folded into the following GEMM: weights from tld_op_ln_fold_weights, row partials from
    tld_op_rowstats_cast, against LayerNorm(x) W^T + bias in fp32."""
    L = lib.load()
    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = torch.randn(M, K, device="cuda", generator=g) * 1.5 + 0.4 * torch.randn(M, 1, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    gamma = 1 + 0.3 * torch.randn(K, device="cuda", generator=g)
    beta = 0.2 * torch.randn(K, device="cuda", generator=g)
    bias = 0.3 * torch.randn(N, device="cuda", generator=g)
    ref = torch.nn.functional.layer_norm(x, (K,), gamma, beta, 1e-5) @ W.t() + bias
    Wf = torch.empty(N, K, device="cuda", dtype=torch.bfloat16)
    s, c = torch.empty(N, device="cuda"), torch.empty(N, device="cuda")
    lib.check(L.tld_op_ln_fold_weights(lib.ptr(W), lib.ptr(gamma), lib.ptr(beta), lib.ptr(bias), lib.ptr(Wf), lib.ptr(s), lib.ptr(c),
                                       N, K, _stream()), "fold")
    assert torch.equal(Wf, (W * gamma).bfloat16())
    assert torch.allclose(s, Wf.float().sum(1), rtol=1e-5, atol=1e-5) and torch.allclose(c, W @ beta + bias, rtol=1e-4, atol=1e-5)
    xb = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
    part = torch.empty(M, K // 32, 2, device="cuda")
    if K % 128 == 0:
        lib.check(L.tld_op_rowstats_cast(lib.ptr(x), lib.ptr(xb), lib.ptr(part), M, K, _stream()), "rowstats")
        assert torch.equal(xb, x.bfloat16())
        assert torch.allclose(part, _row_partials(x), rtol=1e-5, atol=1e-5)
    else:
        xb, part = x.bfloat16(), _row_partials(x)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    lib.check(L.tld_op_gemm_lnfold(lib.ptr(xb), lib.ptr(Wf), M, N, K, lib.ptr(out), lib.ptr(c), lib.ptr(s), lib.ptr(part), K // 32,
                                   _stream()), "gemm_lnfold")
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert rel_fro(out.float(), ref) < 8e-3, _err_map(out.float(), ref)


@pytest.mark.parametrize("M,N,K", [(256, 768, 3072), (4096, 256, 1024), (300, 192, 128), (8192, 768, 512)])
def test_gemm_bias_residual_producer(lib, ctas, M, N, K):
    """mlp.3 + residual with the explicit read-modify-write epilogue that also emits bf16(x_new) and the per-32-column row
    statistics of x_new (the inputs of the next block's folded norm1)."""
    g = torch.Generator(device="cuda").manual_seed(M * 3 + N)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g)
    ref = x + A.float() @ W.float().t() + bias
    xb = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    part = torch.full((M, N // 32, 2), float("nan"), device="cuda")
    lib.check(lib.load().tld_op_gemm_bias_resid_lnp(lib.ptr(A), lib.ptr(W), M, N, K, lib.ptr(x), lib.ptr(bias), lib.ptr(xb),
                                                    lib.ptr(part), _stream()), "gemm_lnp")
    torch.cuda.synchronize()
    assert rel_fro(x, ref) < 2e-5, _err_map(x, ref)
    assert torch.equal(xb, x.bfloat16())                       # the bf16 copy is the rounding of exactly what was stored
    assert torch.allclose(part, _row_partials(x), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("B,n_tok,D", [(2, 64, 128), (3, 256, 768), (16, 256, 768), (1, 1024, 256)])
def test_gemm_cross_attention_producer(lib, ctas, B, n_tok, D):
    """q_linear + 2-key SDPA + residual as an explicit read-modify-write that also emits bf16(x_new) + row partials"""
    g = torch.Generator(device="cuda").manual_seed(13)
    M = B * n_tok
    A = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    Wq = (torch.randn(D, D, device="cuda", generator=g) / math.sqrt(D)).bfloat16()
    kv0 = torch.randn(B, 2 * D, device="cuda", generator=g)
    kv1 = torch.randn(B, 2 * D, device="cuda", generator=g)
    x = torch.randn(M, D, device="cuda", generator=g)
    H = D // 64
    q = (A.float() @ Wq.float().t()).view(B, n_tok, H, 64)
    k = torch.stack([kv0[:, :D], kv1[:, :D]], 1).view(B, 2, H, 64)
    v = torch.stack([kv0[:, D:], kv1[:, D:]], 1).view(B, 2, H, 64)
    sc = torch.einsum("bnhd,bshd->bhns", q, k) / 8.0
    ref = x + torch.einsum("bhns,bshd->bnhd", torch.softmax(sc, -1), v).reshape(M, D)
    xb = torch.full((M, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    part = torch.full((M, D // 32, 2), float("nan"), device="cuda")
    lib.check(lib.load().tld_op_gemm_xattn_lnp(lib.ptr(A), lib.ptr(Wq), M, D, lib.ptr(x), lib.ptr(kv0), lib.ptr(kv1), n_tok,
                                               lib.ptr(xb), lib.ptr(part), _stream()), "gemm_xattn_lnp")
    torch.cuda.synchronize()
    assert rel_fro(x, ref) < 1e-4, _err_map(x, ref)
    assert torch.equal(xb, x.bfloat16())
    assert torch.allclose(part, _row_partials(x), rtol=1e-4, atol=1e-4)
