#!/usr/bin/env python3
"""Time the fused up-projection + dwconv + GELU kernel against the three-kernel path, alone, L2-cold (rotating buffers)."""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from transformer_latent_diffusion_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--reps", type=int, default=12)
ap.add_argument("--mode", default="all")
a = ap.parse_args()
L = _lib.load()
B, K, N = a.batch, 768, 3072
M = B * 256
nb = 3
A = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(nb)]
W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
bias = torch.randn(N, device="cuda")
w9 = torch.randn(9, N, device="cuda") / 3
dwb = torch.randn(N, device="cuda") * 0.1
hid = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(nb)]
out = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(nb)]
st = torch.cuda.current_stream().cuda_stream


def fused(i):
    _lib.check(L.tld_op_gemm_up_dwconv_gelu(A[i].data_ptr(), W.data_ptr(), bias.data_ptr(), None, None, w9.data_ptr(),
                                            dwb.data_ptr(), out[i].data_ptr(), B, K, N, st), "fused")


def gemm(i):
    _lib.check(L.tld_op_gemm(1, A[i].data_ptr(), W.data_ptr(), M, N, K, hid[i].data_ptr(), bias.data_ptr(), st), "gemm")


def conv(i):
    _lib.check(L.tld_op_dwconv_gelu(hid[i].data_ptr(), w9.data_ptr(), dwb.data_ptr(), out[i].data_ptr(), B, 16, N, st), "conv")


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
    timeit(fused, "fused up+dwconv+gelu")
if a.mode in ("all", "split"):
    timeit(gemm, "up-projection GEMM (bias, bf16)")
    timeit(conv, "dwconv+gelu g16")
