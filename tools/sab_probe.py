#!/usr/bin/env python3
"""Time the self-attention backward (all kernels of one call) for both implementations."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from transformer_latent_diffusion_b200 import _lib
L = _lib.load()
for B, n_tok in ((32, 256), (256, 256), (8, 1024), (2, 4096)):
    D = 768
    T = B * n_tok
    qkv = torch.randn(T, 3 * D, device="cuda").bfloat16()
    go = torch.randn(T, D, device="cuda")
    xb = torch.randn(T, D, device="cuda")
    xa = xb + torch.randn(T, D, device="cuda")
    dqkv = torch.empty(T, 3 * D, device="cuda", dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    for impl in (1, 2):
        _lib.check(L.tld_set_option(b"attention_bwd_impl", impl), "opt")
        for _ in range(3):
            _lib.check(L.tld_bwd_self_attention(qkv.data_ptr(), go.data_ptr(), xb.data_ptr(), xa.data_ptr(), dqkv.data_ptr(), B, n_tok, D, st), "sab")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            _lib.check(L.tld_bwd_self_attention(qkv.data_ptr(), go.data_ptr(), xb.data_ptr(), xa.data_ptr(), dqkv.data_ptr(), B, n_tok, D, st), "sab")
        e1.record(); torch.cuda.synchronize()
        print(f"B={B} n_tok={n_tok} impl={impl} ({'mma.sync' if impl == 1 else 'tcgen05 + stats + cast'}): {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call")
