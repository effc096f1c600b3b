#!/usr/bin/env python3
"""Device timing of one training step of the 100M denoiser (fwd + MSE + bwd + all-reduce + Adam + EMA)."""
import argparse, os, sys, copy
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-overlap", action="store_true", help="all-reduce after the backward instead of overlapped with it")
    ap.add_argument("--attn-bwd-impl", type=int, default=0)
    ap.add_argument("--torch-adam", action="store_true", help="torch.optim.Adam(fused=True) + update_ema instead of FusedAdamEMA")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from transformer_latent_diffusion_b200.denoiser import Denoiser
    from transformer_latent_diffusion_b200.train import train_step, update_ema, noise_batch
    from transformer_latent_diffusion_b200 import _lib
    _lib.check(_lib.load().tld_set_option(b"attention_bwd_impl", a.attn_bwd_impl), "opt")
    torch.manual_seed(0)
    m = Denoiser(32, 256, 2, 768, 0, 12).cuda().train()
    m.overlap_grad_allreduce = not a.no_overlap
    ema = copy.deepcopy(m)
    if a.torch_adam:
        opt = torch.optim.Adam(m.parameters(), lr=3e-4, fused=True)
    else:
        from transformer_latent_diffusion_b200.optim import FusedAdamEMA
        opt = FusedAdamEMA(m, lr=3e-4, ema_model=ema, alpha=0.999)
        m.grad_views = True
    B = a.batch
    g = torch.Generator(device="cuda").manual_seed(local)
    x = torch.randn(B, 4, 32, 32, device="cuda", generator=g) * 8
    y = torch.randn(B, 768, device="cuda", generator=g)
    def one():
        sigma = torch.rand(B, device="cuda", dtype=torch.float64)
        eps = torch.randn(B, 4, 32, 32, device="cuda")
        mask = torch.rand(B, device="cuda") < 0.15
        xs, xn, sg, lab = noise_batch(x, y, sigma, eps, mask, 8.0)
        loss = train_step(m, opt, xs, xn, sg, lab)
        if a.torch_adam:
            update_ema(ema, m, 0.999)
        return loss
    for _ in range(3):
        l = one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(a.steps):
        l = one()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    if int(os.environ.get("RANK", "0")) == 0:
        fl = 3 * 46.163e9 * B
        chk = sum(float(p.detach().double().abs().sum()) for p in m.parameters())
        mode = "" if world == 1 else (" all-reduce after backward" if a.no_overlap else " all-reduce overlapped")
        print(f"train step B={B}/GPU x{world}{mode}: {ms:.2f} ms  loss {float(l):.4f}  -> {B * world / ms * 1e3:.0f} samples/s, {fl / ms / 1e9:.0f} TFLOP/s per GPU (3x fwd FLOPs)  param checksum {chk:.6f}")
    if world > 1:
        dist.destroy_process_group()
main()
