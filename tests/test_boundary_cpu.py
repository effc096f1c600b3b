"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
the Python mirror has the reference's constructor signature / attributes / state_dict inventory, and the product
path refuses to run without CUDA (no fallback)."""
import json
import os
import re

import pytest
import torch

from conftest import GOLDEN, ROOT


@pytest.fixture(scope="module")
def built_lib():
    from transformer_latent_diffusion_b200 import build

    return build.build()


def test_library_exports_every_header_symbol(built_lib):
    import ctypes

    from transformer_latent_diffusion_b200 import _lib

    header = open(os.path.join(ROOT, "include", "tld_b200.h")).read()
    declared = set(re.findall(r"TLD_API[^;(]*?\b(tld_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 14
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tld_b200.h but not exported"
    assert declared == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with the header"
    assert _lib.load().tld_version() == 1


def test_no_cuda_means_loud_failure(built_lib):
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import ctypes as C

    from transformer_latent_diffusion_b200 import _lib

    lib = _lib.load()
    cfg = _lib.TldConfig(16, 256, 2, 128, 3, 768, 4, 4, 0.0)
    h = C.c_void_p()
    assert lib.tld_denoiser_create(C.byref(cfg), 0, C.byref(h)) != 0
    assert b"no CUDA device" in lib.tld_last_error()


def test_set_option_validates_keys_and_values(built_lib):
    """tld_set_option is host-only: every documented key is accepted with its documented values, anything else is an error
    with a message (no silent no-op).  Values are restored to the defaults the header documents."""
    from transformer_latent_diffusion_b200 import _lib

    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "tld_b200.h")).read()
    good = {b"gemm_ctas": (0, 1, 2, 0), b"attention_impl": (1, 3, 0), b"attention_bwd_impl": (2, 1, 0), b"fused_mlp": (0, 1),
            b"fused_qkv": (0, 1), b"fused_xattn": (0, 1), b"share_cfg_prefix": (0, 1), b"ln_fold": (1, 0), b"pdl": (1, 0),
            b"qkv_exp_emu": (0, 4, 8, 6), b"xattn_rows": (2, 4), b"xattn_ctas": (2, 1), b"xattn_mma": (1, 2, 3, 0)}
    for key, values in good.items():
        assert key.decode() in header, f"option {key!r} is not documented in include/tld_b200.h"
        for v in values:
            assert lib.tld_set_option(key, v) == 0, (key, v, lib.tld_last_error())
    bad = [(b"xattn_rows", 3), (b"xattn_ctas", 0), (b"xattn_mma", 4), (b"gemm_ctas", 3), (b"attention_impl", 2), (b"qkv_exp_emu", 5),
           (b"no_such_option", 1)]
    for key, v in bad:
        assert lib.tld_set_option(key, v) != 0, (key, v)
        assert lib.tld_last_error()
    assert lib.tld_set_option(None, 0) != 0


def test_denoiser_signature_and_state_dict():
    from transformer_latent_diffusion_b200.configs import DenoiserConfig
    from transformer_latent_diffusion_b200.denoiser import Denoiser
    from dataclasses import asdict

    m = Denoiser(**asdict(DenoiserConfig()))  # tld/diffusion.py:145 construction pattern
    assert (m.image_size, m.noise_embed_dims, m.embed_dim, m.n_channels) == (16, 256, 128, 4)
    with open(os.path.join(GOLDEN, "state_dict_inventory.json")) as f:
        inv = json.load(f)["default"]
    sd = m.state_dict()
    assert list(sd) == list(inv)
    for k, (shape, dtype) in inv.items():
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == "torch." + dtype, k
    assert sum(p.numel() for p in m.parameters()) == 868_800
    big = Denoiser(image_size=32, noise_embed_dims=256, patch_size=2, embed_dim=768, dropout=0, n_layers=12)
    assert sum(p.numel() for p in big.parameters()) == 101_164_352
    # default init statistics match torch's Linear/Conv/LayerNorm/Embedding defaults
    w = big.state_dict()["denoiser_trans_block.decoder_blocks.0.mlp.mlp.0.weight"]
    assert abs(float(w.abs().max()) - 1 / 768 ** 0.5) < 1e-3
    assert torch.equal(big.state_dict()["norm.weight"], torch.ones(768))
    with pytest.raises(Exception):
        m(torch.zeros(1, 4, 16, 16), torch.zeros(1, 1), torch.zeros(1, 768))  # CPU tensors: no fallback


def test_deepcopy_and_load_state_dict():
    import copy

    from transformer_latent_diffusion_b200.denoiser import Denoiser
    from oracle import tld_oracle as O

    cfg = O.OracleCfg()
    m = Denoiser(16, 256, 2, 128, 0, 3)
    m.load_state_dict(O.synth_state_dict(cfg, 1), strict=True)
    c = copy.deepcopy(m)  # tld/train.py:105 builds the EMA model this way
    for (k1, a), (k2, b) in zip(m.state_dict().items(), c.state_dict().items()):
        assert k1 == k2 and torch.equal(a, b) and a.data_ptr() != b.data_ptr()


def test_configs_match_reference_fields():
    from dataclasses import fields

    from transformer_latent_diffusion_b200 import configs as Cc

    assert [f.name for f in fields(Cc.DenoiserConfig)] == [
        "image_size", "noise_embed_dims", "patch_size", "embed_dim", "dropout", "n_layers", "text_emb_size",
        "n_channels", "mlp_multiplier"]
    t = Cc.TrainConfig()
    assert (t.batch_size, t.lr, t.alpha, t.beta_a, t.beta_b) == (128, 3e-4, 0.999, 1, 2.5)
    cfg = Cc.LTDConfig(denoiser_cfg=Cc.DenoiserConfig(embed_dim=256))
    back = Cc.from_json(Cc.LTDConfig, Cc.to_json(cfg))
    assert back == cfg
    Cc.DenoiserConfig(embed_dim=64).check_b200_support()      # any multiple of 64 (heads = embed_dim // 64)
    with pytest.raises(ValueError):
        Cc.DenoiserConfig(embed_dim=96).check_b200_support()


def test_schedule_matches_oracle():
    from transformer_latent_diffusion_b200.diffusion import noise_schedule
    from oracle import tld_oracle as O

    for n, e in [(35, 1), (15, 1), (40, 1), (7, 2.0), (50, 1)]:
        assert noise_schedule(n, e) == O.noise_schedule(n, e)
    assert noise_schedule(3, 1, [0.5, 0.3, 0.1]) == [0.99, 0.3, 0.1]
