#!/usr/bin/env python3
"""Timing of the VAE decode (library path) under different settings."""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from transformer_latent_diffusion_b200.vae import AutoencoderKLDecoder

def run(dtype, chunk, cl, bench, B=64):
    torch.backends.cudnn.benchmark = bench
    m = AutoencoderKLDecoder(chunk=chunk).cuda().to(dtype).eval()
    if cl:
        m = m.to(memory_format=torch.channels_last)
    z = torch.randn(B, 4, 32, 32, device="cuda")
    for _ in range(2):
        m.decode(z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3):
        m.decode(z)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    fl = AutoencoderKLDecoder.flops_per_image(32) * B
    print(f"dtype={dtype} chunk={chunk} channels_last_weights={cl} cudnn.benchmark={bench}: {ms:.1f} ms -> {fl/ms/1e9:.0f} TFLOP/s", flush=True)

for dtype in (torch.bfloat16, torch.float16):
    for chunk in (8, 16, 32):
        for cl in (True,):
            for bench in (False, True):
                try:
                    run(dtype, chunk, cl, bench)
                except Exception as e:
                    print("fail", dtype, chunk, cl, bench, type(e).__name__, str(e)[:100])
from torch.profiler import profile, ProfilerActivity
m = AutoencoderKLDecoder(chunk=16).cuda().to(torch.bfloat16).eval().to(memory_format=torch.channels_last)
z = torch.randn(16, 4, 32, 32, device="cuda")
m.decode(z); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.decode(z); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
