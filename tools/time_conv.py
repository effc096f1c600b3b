#!/usr/bin/env python3
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from transformer_latent_diffusion_b200 import _lib
L = _lib.load()
st = torch.cuda.current_stream().cuda_stream
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (Cin, Cout, H, B) in [(512, 512, 32, 16), (512, 512, 64, 16), (512, 256, 128, 16), (256, 256, 128, 16), (256, 128, 256, 16), (128, 128, 256, 16), (128, 128, 256, 8)]:
    x = torch.randn(B, Cin, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / (3 * Cin ** 0.5)).bfloat16().contiguous(memory_format=torch.channels_last)
    bias = torch.randn(Cout, device="cuda")
    wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    y = torch.empty(B, Cout, H, H, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    fl = 2.0 * B * H * H * Cin * Cout * 9
    res = {}
    for ctas in (1, 2):
        L.tld_set_option(b"gemm_ctas", ctas)
        ms = t(lambda: _lib.check(L.tld_vae_conv3x3(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, H, Cin, Cout, st), "c"))
        res[f"own_cta{ctas}"] = ms
    ms = t(lambda: torch.nn.functional.conv2d(x, w, None, padding=1))
    res["cudnn"] = ms
    print(f"Cin={Cin} Cout={Cout} H={H} B={B}: " + "  ".join(f"{k} {v:.3f} ms ({fl / v / 1e9:.0f} TF/s)" for k, v in res.items()), flush=True)
