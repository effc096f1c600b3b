#!/usr/bin/env python3
"""Time the fused norm2 + cross-attention + residual + norm3 row kernel against the three kernels it replaces (rotating buffers)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from transformer_latent_diffusion_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--mode", default="all")
a = ap.parse_args()
L = _lib.load()
B, D, n_tok = a.batch, a.dim, 256
T, H = B * n_tok, D // 64
nb = 3
x = [torch.randn(T, D, device="cuda") for _ in range(nb)]
y = [torch.empty(T, D, device="cuda", dtype=torch.bfloat16) for _ in range(nb)]
g2, b2, g3, b3 = (torch.randn(D, device="cuda") for _ in range(4))
wq = (torch.randn(D, D, device="cuda") / D ** 0.5).bfloat16()
kv0, kv1 = torch.randn(B, 2 * D, device="cuda"), torch.randn(B, 2 * D, device="cuda")
uk = torch.empty(2 * B, H, D, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def fused(i):
    _lib.check(L.tld_op_ln_xattn_ln(x[i].data_ptr(), g2.data_ptr(), b2.data_ptr(), g3.data_ptr(), b3.data_ptr(), wq.data_ptr(),
                                    kv0.data_ptr(), kv1.data_ptr(), B, n_tok, D, uk.data_ptr(), y[i].data_ptr(), st), "fused")


def split(i):
    _lib.check(L.tld_op_layernorm(x[i].data_ptr(), g2.data_ptr(), b2.data_ptr(), y[i].data_ptr(), T, D, st), "ln2")
    _lib.check(L.tld_op_gemm_xattn(y[i].data_ptr(), wq.data_ptr(), T, D, x[i].data_ptr(), kv0.data_ptr(), kv1.data_ptr(), n_tok, st), "xattn")
    _lib.check(L.tld_op_layernorm(x[i].data_ptr(), g3.data_ptr(), b3.data_ptr(), y[i].data_ptr(), T, D, st), "ln3")


def timeit(fn, name):
    for i in range(3):
        fn(i % nb)
    torch.cuda.synchronize()
    ts = []
    for r in range(a.reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn(r % nb)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"{name}: median {ts[len(ts) // 2]:.1f} us  min {ts[0]:.1f} us")


if a.mode in ("all", "fused"):
    for rows, ctas in ((4, 1), (4, 2), (2, 1), (2, 2), (4, 1), (4, 2)):   # FFMA row kernel: rows per warp, CTAs per SM
        _lib.check(L.tld_set_option(b"xattn_rows", rows), "opt")
        _lib.check(L.tld_set_option(b"xattn_ctas", ctas), "opt")
        timeit(fused, f"fold keys (2 launches) + fused norm2/xattn/norm3, xattn_rows={rows} xattn_ctas={ctas}")
    for mma in (1, 2, 3):   # tf32 mma.sync dots (x truncated / x split / x and keys split)
        _lib.check(L.tld_set_option(b"xattn_mma", mma), "opt")
        timeit(fused, f"fold keys (2 launches) + fused norm2/xattn/norm3, xattn_mma={mma}")
    _lib.check(L.tld_set_option(b"xattn_mma", 0), "opt")
    _lib.check(L.tld_set_option(b"xattn_rows", 4), "opt")
    _lib.check(L.tld_set_option(b"xattn_ctas", 1), "opt")
if a.mode in ("all", "split"):
    timeit(split, "layernorm + q GEMM with 2-key epilogue + layernorm")
