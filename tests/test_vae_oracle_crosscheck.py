"""The VAE oracle against an INDEPENDENT implementation of the same published architecture.

`diffusers` (the reference's VAE, tld/diffusion.py:91, tld/data.py:35-41) is absent offline, so oracle/vae_oracle.py cannot be
pinned to it.  The image does carry another third-party implementation of the very same network: torchtitan's FLUX
`AutoEncoder` (torchtitan/experiments/flux/model/autoencoder.py), i.e. the CompVis / LDM autoencoder that diffusers'
`AutoencoderKL` is a port of (diffusers ships the key map as `convert_ldm_vae_checkpoint`).  This test builds its Encoder /
Decoder at the SDXL-VAE topology (ch_mult 1-2-4-4, two ResNets per level, 4 latent channels), renames the random weights
from the LDM layout to the diffusers layout with that published key map (restated below), and requires the oracle to
reproduce the module's output in fp32.  It checks the ARCHITECTURE restatement (layer order, GroupNorm epsilon, attention
scale, asymmetric down-sampling pad, nearest up-sampling, shortcut placement); it is not a check against diffusers' code or
the madebyollin/sdxl-vae-fp16-fix weights, which stay unavailable.
"""
import pytest
import torch

from oracle import vae_oracle as VO

ae = pytest.importorskip("torchtitan.experiments.flux.model.autoencoder")

CH, CH_MULT, NRES, ZC = 32, (1, 2, 4, 4), 2, 4


def _resnet_map(src, dst):
    out = {}
    for a, b in (("norm1", "norm1"), ("conv1", "conv1"), ("norm2", "norm2"), ("conv2", "conv2"), ("nin_shortcut", "conv_shortcut")):
        for p in ("weight", "bias"):
            out[f"{src}.{a}.{p}"] = f"{dst}.{b}.{p}"
    return out


def _mid_map(side):
    m = {}
    m.update(_resnet_map("mid.block_1", f"{side}.mid_block.resnets.0"))
    m.update(_resnet_map("mid.block_2", f"{side}.mid_block.resnets.1"))
    for a, b in (("norm", "group_norm"), ("q", "to_q"), ("k", "to_k"), ("v", "to_v"), ("proj_out", "to_out.0")):
        for p in ("weight", "bias"):
            m[f"mid.attn_1.{a}.{p}"] = f"{side}.mid_block.attentions.0.{b}.{p}"
    return m


def _to_diffusers(module, side):
    """LDM-layout state_dict of an Encoder / Decoder -> diffusers AutoencoderKL keys (the published convert_ldm_vae_checkpoint map)"""
    L = len(CH_MULT)
    m = _mid_map(side)
    for p in ("weight", "bias"):
        m[f"conv_in.{p}"] = f"{side}.conv_in.{p}"
        m[f"norm_out.{p}"] = f"{side}.conv_norm_out.{p}"
        m[f"conv_out.{p}"] = f"{side}.conv_out.{p}"
    for i in range(L):
        if side == "decoder":
            for j in range(NRES + 1):
                m.update(_resnet_map(f"up.{i}.block.{j}", f"decoder.up_blocks.{L - 1 - i}.resnets.{j}"))
            for p in ("weight", "bias"):
                m[f"up.{i}.upsample.conv.{p}"] = f"decoder.up_blocks.{L - 1 - i}.upsamplers.0.conv.{p}"
        else:
            for j in range(NRES):
                m.update(_resnet_map(f"down.{i}.block.{j}", f"encoder.down_blocks.{i}.resnets.{j}"))
            for p in ("weight", "bias"):
                m[f"down.{i}.downsample.conv.{p}"] = f"encoder.down_blocks.{i}.downsamplers.0.conv.{p}"
    sd = {}
    for k, v in module.state_dict().items():
        assert k in m, f"unmapped LDM key {k}"
        v = v.detach().clone()
        if ".attentions.0.to_" in m[k] and v.dim() == 4:   # 1x1 conv -> Linear
            v = v[:, :, 0, 0]
        sd[m[k]] = v
    return sd


def _randomize(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("weight") and p.dim() == 1:       # GroupNorm gamma around 1
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            else:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / fan_in ** 0.5)


def test_decoder_oracle_matches_the_ldm_autoencoder():
    dec = ae.Decoder(ch=CH, out_ch=3, ch_mult=list(CH_MULT), num_res_blocks=NRES, in_channels=3, resolution=64, z_channels=ZC).eval()
    _randomize(dec, 11)
    sd = _to_diffusers(dec, "decoder")
    g = torch.Generator().manual_seed(12)
    sd["post_quant_conv.weight"] = torch.randn(ZC, ZC, 1, 1, generator=g) * 0.5
    sd["post_quant_conv.bias"] = torch.randn(ZC, generator=g) * 0.1
    z = torch.randn(2, ZC, 8, 8, generator=g)
    with torch.no_grad():
        ref = dec(torch.nn.functional.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"]))
        out = VO.decode(sd, z)
    assert out.shape == ref.shape == (2, 3, 64, 64)
    err = float((out - ref).norm() / ref.norm())
    assert err < 1e-5, f"decoder restatement vs LDM AutoEncoder: rel_fro={err:.3e}"


def test_encoder_oracle_matches_the_ldm_autoencoder():
    enc = ae.Encoder(resolution=64, in_channels=3, ch=CH, ch_mult=list(CH_MULT), num_res_blocks=NRES, z_channels=ZC).eval()
    _randomize(enc, 21)
    sd = _to_diffusers(enc, "encoder")
    g = torch.Generator().manual_seed(22)
    sd["quant_conv.weight"] = torch.randn(2 * ZC, 2 * ZC, 1, 1, generator=g) * 0.4
    sd["quant_conv.bias"] = torch.randn(2 * ZC, generator=g) * 0.1
    x = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    with torch.no_grad():
        moments = torch.nn.functional.conv2d(enc(x), sd["quant_conv.weight"], sd["quant_conv.bias"])
        mean_ref, logvar_ref = moments.chunk(2, dim=1)
        mean, logvar = VO.encode_moments(sd, x)
    assert mean.shape == (2, ZC, 8, 8)
    assert float((mean - mean_ref).norm() / mean_ref.norm()) < 1e-5
    assert float((logvar - logvar_ref.clamp(-30.0, 20.0)).norm() / logvar_ref.norm()) < 1e-5
