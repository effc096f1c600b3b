"""The oracle (oracle/tld_oracle.py) against the golden outputs of the unmodified reference.

Fixtures were produced by tests/golden/make_golden.py from /root/reference (see manifest.json).
Tolerance: both sides are fp32 on CPU with different op orderings; the fp32-vs-fp64 noise floor
of one forward is rel-Fro 1.4e-5, so the bar here is rel-Fro <= 1e-4.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_fro
from oracle import tld_oracle as O

with open(os.path.join(GOLDEN, "manifest.json")) as f:
    MANIFEST = json.load(f)["cases"]

FWD = [k for k, v in MANIFEST.items() if v["kind"] == "forward"]
GEN = [k for k, v in MANIFEST.items() if v["kind"] == "sampler"]
TOL = 1e-4


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.mark.parametrize("name", FWD)
def test_forward_matches_reference(name):
    meta = MANIFEST[name]
    cfg = O.OracleCfg(**meta["cfg"])
    sd = O.synth_state_dict(cfg, meta["weight_seed"])
    z = _load(name)
    with torch.no_grad():
        out, inter = O.denoiser_forward(sd, cfg, z["x"], z["noise_level"], z["label"], return_intermediates=True)
    assert out.shape == z["out"].shape
    assert rel_fro(inter["block0"], z["block_first"]) < TOL
    assert rel_fro(inter[f"block{cfg.n_layers - 1}"], z["block_last"]) < TOL
    assert rel_fro(out, z["out"]) < TOL


@pytest.mark.parametrize("name", GEN)
def test_sampler_matches_reference(name):
    meta = MANIFEST[name]
    cfg = O.OracleCfg(**meta["cfg"])
    sd = O.synth_state_dict(cfg, meta["weight_seed"])
    z = _load(name)
    trace = []
    with torch.no_grad():
        lat = O.generate_latents(sd, cfg, z["labels"], z["seeds"], trace=trace, **meta["kwargs"])
    # the same noise levels reach the model, in the same order, as in the reference loop
    np.testing.assert_allclose([s for _, s, _ in trace], z["sigmas_seen"].numpy(), rtol=1e-6)
    assert len(trace) == meta["kwargs"]["n_iter"] + (1 if "noise_levels" in meta["kwargs"] else 0)
    assert rel_fro(lat, z["latent"]) < TOL
    # IdentityVAE: img == latent * scale_factor (tld/diffusion.py:91)
    assert rel_fro(lat * 8, z["img"]) < TOL


def test_state_dict_inventory_matches_reference():
    with open(os.path.join(GOLDEN, "state_dict_inventory.json")) as f:
        inv = json.load(f)
    for name, cfg in {"default": O.OracleCfg(),
                      "100m_256px": O.OracleCfg(image_size=32, embed_dim=768, n_layers=12)}.items():
        mine = {k: list(v) for k, v in O.param_shapes(cfg).items()}
        ref = {k: v[0] for k, v in inv[name].items()}
        assert mine == ref
        assert list(mine) == list(ref)  # same order as the reference's state_dict
    assert inv["100m_256px_n_params"] == 101_164_352


def test_train_glue_matches_reference():
    meta = MANIFEST["train_glue"]
    cfg = O.OracleCfg(**meta["cfg"])
    sd = O.synth_state_dict(cfg, meta["weight_seed"])
    z = _load("train_glue")
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and not k.endswith("angular_speeds")}
    live = dict(sd)
    live.update(params)
    ema = {k: v.detach().clone() for k, v in live.items()}
    opt = torch.optim.Adam(list(params.values()), lr=meta["lr"])
    losses = []
    for step in range(2):
        x, x_noisy, t, lab = O.noise_inputs(z[f"x{step}"], z[f"sigma{step}"], z[f"eps{step}"],
                                            z[f"y{step}"], z[f"mask{step}"])
        opt.zero_grad()
        pred = O.denoiser_forward(live, cfg, x_noisy, t, lab)
        loss = torch.nn.functional.mse_loss(pred, x)
        loss.backward()
        opt.step()
        with torch.no_grad():
            O.ema_update(ema, {k: v.detach() for k, v in live.items()}, meta["alpha"])
        losses.append(float(loss.detach()))
    np.testing.assert_allclose(losses, z["losses"].numpy(), rtol=2e-5)
    for k in meta["probe"]:
        # Adam's first steps move every weight by ~lr regardless of gradient scale, so compare the
        # update (after - before), not the weight, to make the check meaningful.
        d_mine = live[k].detach() - sd[k]
        d_ref = z["after." + k] - sd[k]
        assert rel_fro(d_mine, d_ref) < 2e-2, k
        assert rel_fro(ema[k], z["ema." + k]) < 1e-6, k


def test_flops_formula():
    # BASELINE.md §3 values
    assert abs(O.flops_forward(O.OracleCfg(image_size=32, embed_dim=768, n_layers=12)) / 1e9 - 46.163) < 1e-2
    assert abs(O.flops_forward(O.OracleCfg(image_size=64, embed_dim=768, n_layers=12)) / 1e9 - 213.466) < 1e-2
    assert abs(O.flops_forward(O.OracleCfg(image_size=32, embed_dim=256, n_layers=3)) / 1e9 - 1.432) < 1e-2
