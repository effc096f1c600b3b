#!/usr/bin/env python3
"""Developer tool: timeline of CTA 0 of the fused qkv + attention kernel (library built with TLD_NVCC_EXTRA=-DTLD_TRACE)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from transformer_latent_diffusion_b200 import _lib

L = _lib.load()
raw = C.CDLL(_lib.LIB_PATH) if hasattr(_lib, "LIB_PATH") else L
B, D = 128, 768
xn = torch.randn(B * 256, D, device="cuda").bfloat16()
W = (torch.randn(3 * D, D, device="cuda") / D ** 0.5).bfloat16()
x = torch.randn(B * 256, D, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(L.tld_op_qkv_attention(xn.data_ptr(), W.data_ptr(), x.data_ptr(), B, 256, D, st), "fused")
torch.cuda.synchronize()
buf = (C.c_ulonglong * (4 * 2048))()
fn = raw.tld_debug_qa_trace
fn.argtypes = [C.c_void_p]
assert fn(buf) == 0
NAMES = {0: {1: "G first stage full", 2: "G issued", 3: "p_full(i-1)", 4: "qk_ready"},
         1: {1: "acc_full", 2: "drain done", 3: "epilogue done", 4: "s_full", 5: "P written", 6: "qk stored", 7: "o_full(i-1)"},
         2: {1: "drain_done barrier", 2: "fence done"},
         3: {1: "acc_full", 2: "drain done", 3: "epilogue done", 4: "s_full", 5: "P written", 6: "qk stored", 7: "o_full(i-1)"}}
ROLE = ["MMA ", "SMAX0", "FWD ", "SMAX1"]
ev = []
for role in range(4):
    for i in range(2048):
        v = buf[role * 2048 + i]
        if v == 0:
            break
        ev.append((v >> 8, role, v & 255))
ev.sort()
t0 = ev[0][0]
# items 8..11 (steady state)
cnt = 0
last = {}
for t, role, ident in ev:
    if role == 0 and ident == 1:
        cnt += 1
    if 8 <= cnt <= 11:
        print(f"{t - t0:8d}  {ROLE[role]}  {NAMES[role][ident]}")
