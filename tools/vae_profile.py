"""Kernel-level breakdown of one VAE decode of 64 latents (torch profiler, CUDA activities)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from torch.profiler import profile, ProfilerActivity
from transformer_latent_diffusion_b200.vae import AutoencoderKLDecoder
m = AutoencoderKLDecoder().cuda().to(torch.bfloat16).eval()
z = torch.randn(64, 4, 32, 32, device="cuda")
for _ in range(2):
    m.decode(z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); m.decode(z); e1.record(); torch.cuda.synchronize()
print(f"decode 64 latents: {e0.elapsed_time(e1):.1f} ms; own launches {m.own_launches}")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.decode(z); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))
