#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Weights come from `oracle.tld_oracle.synth_state_dict(cfg, seed)` (deterministic torch CPU stream),
loaded into the reference `tld.denoiser.Denoiser` with a strict `load_state_dict`, so the fixture
stores only the seed, the inputs and the reference's outputs.
"""
from __future__ import annotations

import json
import os
import sys
from dataclasses import asdict

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle import tld_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

FWD_CASES = {
    # name: (cfg, batch, weight seed, input seed)
    "fwd_default": (O.OracleCfg(), 4, 1, 101),  # tests/test_diffuser.py:27-46 shape
    "fwd_cfg1": (O.OracleCfg(image_size=32, embed_dim=256, n_layers=3), 2, 2, 102),  # BASELINE cfg[0]
    "fwd_d64_odd": (O.OracleCfg(image_size=16, embed_dim=64, n_layers=1), 3, 3, 103),
    "fwd_d768_l1": (O.OracleCfg(image_size=16, embed_dim=768, n_layers=1), 2, 4, 104),
}

SAMPLER_CASES = {
    # name: (cfg, num_imgs, wseed, kwargs)
    "gen_dpmpp": (O.OracleCfg(image_size=16, embed_dim=128, n_layers=2), 2, 11,
                  dict(n_iter=6, class_guidance=3.0, exponent=1, sharp_f=0.1, bright_f=0.1,
                       use_ddpm_plus=True)),
    "gen_ddim_exp2": (O.OracleCfg(image_size=16, embed_dim=128, n_layers=2), 3, 12,
                      dict(n_iter=5, class_guidance=6.0, exponent=2, sharp_f=0.0, bright_f=0.0,
                           use_ddpm_plus=False)),
    "gen_custom_levels": (O.OracleCfg(image_size=16, embed_dim=128, n_layers=1), 1, 13,
                          dict(n_iter=4, class_guidance=1.5, exponent=1, sharp_f=0.0, bright_f=0.2,
                               use_ddpm_plus=True, noise_levels=[0.9, 0.7, 0.45, 0.2, 0.05])),
}


def ref_model(ref, cfg: O.OracleCfg, wseed: int):
    m = ref.denoiser.Denoiser(**asdict(cfg))
    sd = O.synth_state_dict(cfg, wseed)
    m.load_state_dict(sd, strict=True)
    return m.eval(), sd


def inputs(cfg: O.OracleCfg, B: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, cfg.n_channels, cfg.image_size, cfg.image_size, generator=g)
    t = torch.rand(B, 1, generator=g)
    lab = torch.randn(B, cfg.text_emb_size, generator=g)
    return x, t, lab


def main():
    ref = ref_loader.load_reference()
    torch.set_grad_enabled(False)
    manifest = {"reference_root": ref_loader.REFERENCE_ROOT, "torch": torch.__version__, "cases": {}}

    # state_dict inventory of the reference (keys, shapes, dtypes) for the 100M and default configs
    inv = {}
    for name, cfg in {"default": O.OracleCfg(),
                      "100m_256px": O.OracleCfg(image_size=32, embed_dim=768, n_layers=12)}.items():
        m = ref.denoiser.Denoiser(**asdict(cfg))
        inv[name] = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in m.state_dict().items()}
        inv[name + "_n_params"] = sum(p.numel() for p in m.parameters())
    with open(os.path.join(OUT, "state_dict_inventory.json"), "w") as f:
        json.dump(inv, f, indent=0, sort_keys=False)

    for name, (cfg, B, wseed, iseed) in FWD_CASES.items():
        m, _ = ref_model(ref, cfg, wseed)
        x, t, lab = inputs(cfg, B, iseed)
        blocks = []
        hooks = [blk.register_forward_hook(lambda mod, a, out: blocks.append(out.clone()))
                 for blk in m.denoiser_trans_block.decoder_blocks]
        out = m(x, t, lab)
        for h in hooks:
            h.remove()
        arrs = dict(x=x.numpy(), noise_level=t.numpy(), label=lab.numpy(), out=out.numpy(),
                    block_last=blocks[-1].numpy(), block_first=blocks[0].numpy())
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
        manifest["cases"][name] = dict(kind="forward", cfg=asdict(cfg), batch=B, weight_seed=wseed,
                                       input_seed=iseed)
        print(name, tuple(out.shape), float(out.abs().mean()))

    for name, (cfg, num, wseed, kw) in SAMPLER_CASES.items():
        m, _ = ref_model(ref, cfg, wseed)
        g = torch.Generator().manual_seed(1000 + wseed)
        labels = torch.randn(num, cfg.text_emb_size, generator=g)
        seeds = torch.randn(num, cfg.n_channels, cfg.image_size, cfg.image_size, generator=g)
        calls = []

        class Spy(torch.nn.Module):  # records what the reference loop feeds the model
            def __init__(self, inner):
                super().__init__()
                self.inner = inner
                self.n_channels = inner.n_channels

            def forward(self, xx, tt, ll):
                calls.append(float(tt[0, 0]))
                return self.inner(xx, tt, ll)

        gen = ref.diffusion.DiffusionGenerator(Spy(m), ref_loader.IdentityVAE(), torch.device("cpu"),
                                               torch.float32)
        kw2 = dict(kw)
        if "noise_levels" in kw2:
            kw2["noise_levels"] = list(kw2["noise_levels"])
        img, latent = gen.generate(labels=labels, num_imgs=num, img_size=cfg.image_size, seeds=seeds,
                                   scale_factor=8, **kw2)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), labels=labels.numpy(), seeds=seeds.numpy(),
                            latent=latent.numpy(), img=img.numpy(), sigmas_seen=np.array(calls))
        manifest["cases"][name] = dict(kind="sampler", cfg=asdict(cfg), num_imgs=num, weight_seed=wseed,
                                       kwargs=kw)
        print(name, len(calls), "model calls", float(latent.abs().mean()))

    # train-step glue: 2 Adam steps + EMA on the reference module with injected randomness
    torch.set_grad_enabled(True)
    cfg = O.OracleCfg(image_size=16, embed_dim=64, n_layers=1)
    m, sd0 = ref_model(ref, cfg, 21)
    m.train()
    import copy
    ema = copy.deepcopy(m)
    opt = torch.optim.Adam(m.parameters(), lr=3e-4)
    g = torch.Generator().manual_seed(77)
    B = 6
    losses, rec = [], {}
    for step in range(2):
        x = torch.randn(B, 4, 16, 16, generator=g) * 8
        y = torch.randn(B, 768, generator=g)
        sigma = torch.rand(B, generator=g, dtype=torch.float64)
        eps = torch.randn(B, 4, 16, 16, generator=g)
        mask = torch.rand(B, generator=g) < 0.3
        rec[f"x{step}"], rec[f"y{step}"], rec[f"sigma{step}"] = x.numpy(), y.numpy(), sigma.numpy()
        rec[f"eps{step}"], rec[f"mask{step}"] = eps.numpy(), mask.numpy()
        # tld/train.py:122-138
        xs = x / 8
        x_noisy = (sigma.view(-1, 1, 1, 1) * eps + (1 - sigma).view(-1, 1, 1, 1) * xs).float()
        label = y.clone()
        label[mask] = 0
        opt.zero_grad()
        pred = m(x_noisy, sigma.float().view(-1, 1), label)
        loss = torch.nn.functional.mse_loss(pred, xs)
        loss.backward()
        opt.step()
        with torch.no_grad():  # tld/train.py:55-58
            for pe, pm in zip(ema.parameters(), m.parameters()):
                pe.data.mul_(0.999).add_(pm.data, alpha=0.001)
        losses.append(float(loss))
    probe = ["label_proj.weight", "denoiser_trans_block.decoder_blocks.0.mlp.mlp.1.weight",
             "denoiser_trans_block.decoder_blocks.0.self_attention.qkv_linear.weight",
             "denoiser_trans_block.patchify_and_embed.0.weight", "fourier_feats.1.bias"]
    for k in probe:
        rec["after." + k] = m.state_dict()[k].detach().numpy()
        rec["ema." + k] = ema.state_dict()[k].detach().numpy()
    rec["losses"] = np.array(losses)
    np.savez_compressed(os.path.join(OUT, "train_glue.npz"), **rec)
    manifest["cases"]["train_glue"] = dict(kind="train", cfg=asdict(cfg), weight_seed=21, batch=B,
                                           lr=3e-4, alpha=0.999, probe=probe)
    print("train_glue losses", losses)

    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
