"""Backward-pass kernels (training step) against torch autograd of the same op in fp32, same bf16-rounded inputs."""
import math

import pytest
import torch

from conftest import rel_fro

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from transformer_latent_diffusion_b200 import _lib

    return _lib


def _st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("R,C", [(64, 64), (200, 768), (1000, 3072), (8192, 256)])
def test_cast_transpose_and_colsum(lib, R, C):
    g = torch.Generator(device="cuda").manual_seed(R + C)
    x = torch.randn(R, C, device="cuda", generator=g)
    o = torch.empty(R, C, device="cuda", dtype=torch.bfloat16)
    ot = torch.empty(C, R, device="cuda", dtype=torch.bfloat16)
    lib.check(lib.load().tld_bwd_cast_transpose(lib.ptr(x), lib.ptr(o), lib.ptr(ot), R, C, _st()), "ct")
    assert torch.equal(o, x.bfloat16()) and torch.equal(ot, x.bfloat16().t().contiguous())
    s = torch.empty(C, device="cuda")
    lib.check(lib.load().tld_bwd_colsum(lib.ptr(x), lib.ptr(s), R, C, _st()), "cs")
    assert rel_fro(s, x.double().sum(0)) < 1e-5


@pytest.mark.parametrize("rows,D", [(8, 128), (77, 256), (2048, 768)])
def test_layernorm_backward(lib, rows, D):
    g = torch.Generator(device="cuda").manual_seed(D)
    x = (torch.randn(rows, D, device="cuda", generator=g) * 2 + 0.3).requires_grad_(True)
    w = torch.randn(D, device="cuda", generator=g).requires_grad_(True)
    b = torch.randn(D, device="cuda", generator=g).requires_grad_(True)
    dy = torch.randn(rows, D, device="cuda", generator=g)
    y = torch.nn.functional.layer_norm(x, (D,), w, b, 1e-5)
    y.backward(dy)
    dx0 = torch.randn(rows, D, device="cuda", generator=g)  # the kernel accumulates into dx
    dx = dx0.clone()
    dw, db = torch.empty(D, device="cuda"), torch.empty(D, device="cuda")
    lib.check(lib.load().tld_bwd_layernorm(lib.ptr(dy), lib.ptr(x.detach()), lib.ptr(w.detach()), lib.ptr(dx), lib.ptr(dw),
                                           lib.ptr(db), rows, D, _st()), "lnb")
    assert rel_fro(dx - dx0, x.grad) < 1e-4
    assert rel_fro(dw, w.grad) < 1e-4 and rel_fro(db, b.grad) < 1e-4


@pytest.mark.parametrize("B,G,C", [(2, 8, 512), (3, 16, 1024), (2, 16, 3072), (1, 32, 256)])
def test_dwconv_gelu_backward(lib, B, G, C):
    g = torch.Generator(device="cuda").manual_seed(G + C)
    h = torch.randn(B, G, G, C, device="cuda", generator=g).bfloat16()
    dg = torch.randn(B, G, G, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(C, 1, 3, 3, device="cuda", generator=g) / 3).requires_grad_(True)
    bias = (torch.randn(C, device="cuda", generator=g) * 0.1).requires_grad_(True)
    hf = h.float().permute(0, 3, 1, 2).requires_grad_(True)
    out = torch.nn.functional.gelu(torch.nn.functional.conv2d(hf, w, bias, padding=1, groups=C))
    out.backward(dg.float().permute(0, 3, 1, 2))
    w9 = w.detach().view(C, 9).t().contiguous()
    du = torch.empty_like(h)
    dh = torch.empty_like(h)
    dw9 = torch.empty(9, C, device="cuda")
    db = torch.empty(C, device="cuda")
    lib.check(lib.load().tld_bwd_dwconv_gelu(lib.ptr(h), lib.ptr(dg), lib.ptr(w9), lib.ptr(bias.detach()), lib.ptr(du), lib.ptr(dh),
                                             lib.ptr(dw9), lib.ptr(db), B, G, C, _st()), "dwb")
    # du is rounded to bf16 between the two conv passes: ~2^-9 relative per element
    assert rel_fro(dh.float(), hf.grad.permute(0, 2, 3, 1)) < 8e-3
    assert rel_fro(dw9, w.grad.view(C, 9).t()) < 8e-3
    assert rel_fro(db, bias.grad) < 8e-3


@pytest.mark.parametrize("B,n_tok,D", [(2, 64, 128), (3, 256, 768), (1, 1024, 128)])
def test_cross_attention_backward(lib, B, n_tok, D):
    g = torch.Generator(device="cuda").manual_seed(n_tok + D)
    T, H = B * n_tok, D // 64
    q = torch.randn(T, D, device="cuda", generator=g).bfloat16()
    go = torch.randn(T, D, device="cuda", generator=g)
    kv0 = torch.randn(B, 2 * D, device="cuda", generator=g).requires_grad_(True)
    kv1 = torch.randn(B, 2 * D, device="cuda", generator=g).requires_grad_(True)
    qf = q.float().requires_grad_(True)
    qh = qf.view(B, n_tok, H, 64)
    k = torch.stack([kv0[:, :D], kv1[:, :D]], 1).view(B, 2, H, 64)
    v = torch.stack([kv0[:, D:], kv1[:, D:]], 1).view(B, 2, H, 64)
    s = torch.einsum("bnhd,bshd->bhns", qh, k) / 8.0
    o = torch.einsum("bhns,bshd->bnhd", torch.softmax(s, -1), v).reshape(T, D)
    o.backward(go)
    dq = torch.empty(T, D, device="cuda", dtype=torch.bfloat16)
    d0 = torch.zeros(B, 2 * D, device="cuda")
    d1 = torch.zeros(B, 2 * D, device="cuda")
    lib.check(lib.load().tld_bwd_xattn(lib.ptr(q), lib.ptr(go), lib.ptr(kv0.detach()), lib.ptr(kv1.detach()), lib.ptr(dq), lib.ptr(d0),
                                       lib.ptr(d1), B, n_tok, D, _st()), "xb")
    assert rel_fro(dq.float(), qf.grad) < 5e-3
    assert rel_fro(d0, kv0.grad) < 1e-4 and rel_fro(d1, kv1.grad) < 1e-4


@pytest.mark.parametrize("B,n_tok,D", [(1, 64, 128), (2, 256, 128), (2, 256, 768), (3, 128, 256),
                                       (2, 512, 128), (1, 1024, 192), (1, 4096, 64),    # > 256 tokens: the key-tiled kernels
                                       (40, 256, 768)])                                 # several CTAs per SM in sequence
@pytest.mark.parametrize("impl", [1, 2])      # 1: mma.sync kernels (default), 2: tcgen05 kernel where it applies (tokens % 256 == 0)
def test_self_attention_backward(lib, B, n_tok, D, impl):
    lib.check(lib.load().tld_set_option(b"attention_bwd_impl", impl), "opt")
    try:
        _self_attention_backward(lib, B, n_tok, D)
    finally:
        lib.check(lib.load().tld_set_option(b"attention_bwd_impl", 0), "opt")


def _self_attention_backward(lib, B, n_tok, D):
    g = torch.Generator(device="cuda").manual_seed(n_tok * 3 + D)
    T, H = B * n_tok, D // 64
    qkv = torch.randn(T, 3 * D, device="cuda", generator=g).bfloat16()
    go = torch.randn(T, D, device="cuda", generator=g)
    f = qkv.float().requires_grad_(True)
    qh, kh, vh = (t.view(B, n_tok, H, 64).permute(0, 2, 1, 3) for t in f.chunk(3, dim=1))
    o = (torch.softmax(qh @ kh.transpose(-1, -2) / 8.0, -1) @ vh).permute(0, 2, 1, 3).reshape(T, D)
    o.backward(go)
    xb = torch.randn(T, D, device="cuda", generator=g)
    xa = xb + o.detach()
    dqkv = torch.full((T, 3 * D), float("nan"), device="cuda", dtype=torch.bfloat16)
    lib.check(lib.load().tld_bwd_self_attention(lib.ptr(qkv), lib.ptr(go), lib.ptr(xb), lib.ptr(xa), lib.ptr(dqkv), B, n_tok, D,
                                                _st()), "sab")
    torch.cuda.synchronize()
    assert torch.isfinite(dqkv.float()).all()
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        err = rel_fro(dqkv.float()[:, sl], f.grad[:, sl])
        assert err < 1.5e-2, (name, err)


@pytest.mark.parametrize("T,N,K", [(512, 768, 3072), (4096, 2304, 768)])
def test_linear_backward_via_gemm(lib, T, N, K):
    """dgrad and wgrad of y = x W^T through the tcgen05 GEMM on transposed bf16 copies."""
    g = torch.Generator(device="cuda").manual_seed(T + N)
    x = torch.randn(T, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    dy = torch.randn(T, N, device="cuda", generator=g)
    L = lib.load()
    dyb = torch.empty(T, N, device="cuda", dtype=torch.bfloat16)
    dybT = torch.empty(N, T, device="cuda", dtype=torch.bfloat16)
    xbT = torch.empty(K, T, device="cuda", dtype=torch.bfloat16)
    lib.check(L.tld_bwd_cast_transpose(lib.ptr(dy), lib.ptr(dyb), lib.ptr(dybT), T, N, _st()), "ct")
    lib.check(L.tld_bwd_cast_transpose(lib.ptr(x), None, lib.ptr(xbT), T, K, _st()), "ct")
    WT = W.t().contiguous().bfloat16()  # [K, N]
    dx = torch.empty(T, K, device="cuda")
    dW = torch.empty(N, K, device="cuda")
    lib.check(L.tld_op_gemm(4, lib.ptr(dyb), lib.ptr(WT), T, K, N, lib.ptr(dx), None, _st()), "dgrad")   # dx = dy W
    lib.check(L.tld_op_gemm(4, lib.ptr(dybT), lib.ptr(xbT), N, K, T, lib.ptr(dW), None, _st()), "wgrad")  # dW = dy^T x
    assert rel_fro(dx, dy.bfloat16().float() @ W.bfloat16().float()) < 1e-4
    assert rel_fro(dW, dy.bfloat16().float().t() @ x.bfloat16().float()) < 1e-4
