#!/usr/bin/env python3
"""Quick device timing of Denoiser.forward / the sampler at the 100M config (development aid, not the bench)."""
import argparse
import sys
import os
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from transformer_latent_diffusion_b200.denoiser import Denoiser
from transformer_latent_diffusion_b200.diffusion import DiffusionGenerator


class _Id:
    def decode(self, z):
        return (z,)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--img", type=int, default=32)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=35)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--forward-only", action="store_true")
    ap.add_argument("--gemm-ctas", type=int, default=0)
    ap.add_argument("--attn-impl", type=int, default=0)
    ap.add_argument("--exp-emu", type=int, default=-1)
    ap.add_argument("--pdl", type=int, default=0)
    ap.add_argument("--fused-mlp", type=int, default=1)
    ap.add_argument("--fused-qkv", type=int, default=1)
    ap.add_argument("--qkv-emu", type=int, default=6)
    ap.add_argument("--share-cfg", type=int, default=1)
    ap.add_argument("--xattn-mma", type=int, default=0)
    ap.add_argument("--xattn-rows", type=int, default=4)
    ap.add_argument("--xattn-ctas", type=int, default=1)
    a = ap.parse_args()
    from transformer_latent_diffusion_b200 import _lib
    _lib.check(_lib.load().tld_set_option(b"gemm_ctas", a.gemm_ctas), "opt")
    _lib.check(_lib.load().tld_set_option(b"attention_impl", a.attn_impl), "opt")
    if a.exp_emu >= 0:
        _lib.check(_lib.load().tld_set_option(b"attention_exp_emu", a.exp_emu), "opt")
    _lib.check(_lib.load().tld_set_option(b"pdl", a.pdl), "opt")
    _lib.check(_lib.load().tld_set_option(b"fused_mlp", a.fused_mlp), "opt")
    _lib.check(_lib.load().tld_set_option(b"fused_qkv", a.fused_qkv), "opt")
    _lib.check(_lib.load().tld_set_option(b"qkv_exp_emu", a.qkv_emu), "opt")
    _lib.check(_lib.load().tld_set_option(b"share_cfg_prefix", a.share_cfg), "opt")
    _lib.check(_lib.load().tld_set_option(b"xattn_mma", a.xattn_mma), "opt")
    _lib.check(_lib.load().tld_set_option(b"xattn_rows", a.xattn_rows), "opt")
    _lib.check(_lib.load().tld_set_option(b"xattn_ctas", a.xattn_ctas), "opt")
    torch.manual_seed(0)
    m = Denoiser(a.img, 256, 2, 768, 0, 12).cuda().eval()
    B2 = 2 * a.batch
    x = torch.randn(B2, 4, a.img, a.img, device="cuda")
    t = torch.rand(B2, 1, device="cuda")
    lab = torch.randn(B2, 768, device="cuda")
    with torch.no_grad():
        for _ in range(2):
            m(x, t, lab)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(a.reps):
            m(x, t, lab)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    N = (a.img // 2) ** 2
    D, L = 768, 12
    flops = B2 * (L * (24 * N * D * D + 4 * N * N * D + 80 * N * D + 8 * D * D))
    print(f"forward B_eff={B2} img={a.img}: {ms:.3f} ms  -> {flops / ms / 1e9:.1f} TFLOP/s")
    if a.forward_only:
        return
    gen = DiffusionGenerator(m, _Id(), torch.device("cuda:0"), torch.float32)
    labels = torch.randn(a.batch, 768, device="cuda")
    seeds = torch.randn(a.batch, 4, a.img, a.img, device="cuda")
    for r in range(a.reps):
        t0 = time.time()
        gen.generate_latents(labels, n_iter=a.steps, num_imgs=a.batch, img_size=a.img, seeds=seeds, class_guidance=6)
        torch.cuda.synchronize()
        wall = time.time() - t0
        ms, launches = gen.last_stats()
        print(f"sampler B={a.batch} steps={a.steps}: loop {ms:.2f} ms ({ms / a.steps:.3f} ms/step, {launches} launches), "
              f"wall {wall * 1e3:.1f} ms -> {a.batch / (ms / 1e3):.1f} img/s denoiser-only")


if __name__ == "__main__":
    main()
