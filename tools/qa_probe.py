#!/usr/bin/env python3
"""Time the fused qkv-projection + attention kernel against the two-kernel path (qkv GEMM, attention_tc2), alone, rotating buffers."""
import argparse
import math
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
B, D = a.batch, a.dim
M = B * 256
nb = 3
xn = [torch.randn(M, D, device="cuda").bfloat16() for _ in range(nb)]
W = (torch.randn(3 * D, D, device="cuda") / math.sqrt(D)).bfloat16()
x = [torch.randn(M, D, device="cuda") for _ in range(nb)]
qkv = [torch.empty(M, 3 * D, device="cuda", dtype=torch.bfloat16) for _ in range(nb)]
st = torch.cuda.current_stream().cuda_stream


def fused(i):
    _lib.check(L.tld_op_qkv_attention(xn[i].data_ptr(), W.data_ptr(), x[i].data_ptr(), B, 256, D, st), "fused")


def gemm(i):
    _lib.check(L.tld_op_gemm(0, xn[i].data_ptr(), W.data_ptr(), M, 3 * D, D, qkv[i].data_ptr(), None, st), "gemm")


def attn(i):
    _lib.check(L.tld_op_self_attention(qkv[i].data_ptr(), x[i].data_ptr(), B, 256, D, 3, st), "attn")


def both(i):
    gemm(i)
    attn(i)


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
    timeit(fused, "fused qkv + attention")
if a.mode in ("all", "split"):
    timeit(gemm, "qkv GEMM")
    timeit(attn, "attention_tc2")
    timeit(both, "qkv GEMM + attention_tc2")
