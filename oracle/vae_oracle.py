"""fp32 CPU restatement of the SDXL AutoencoderKL *decoder* (third-party: `diffusers`, unpinned in the
reference's pyproject.toml:26; weights `madebyollin/sdxl-vae-fp16-fix`, tld/configs.py:42).  TEST INFRA ONLY.

PARITY NOT PINNED TO DIFFUSERS: neither `diffusers` nor the weights exist in this environment, and the reference's tests
that touch the VAE assert no values.  This file restates the published decoder algorithm
(GroupNorm(32, eps 1e-6) -> SiLU -> conv3x3 ResNet blocks, one single-head attention in the mid block,
nearest-2x upsampling + conv3x3, final GroupNorm/SiLU/conv) directly from tensors in a diffusers-keyed
state_dict, written independently of transformer_latent_diffusion_b200/vae.py (no nn.Module, explicit
GroupNorm/softmax) so the two can be compared on random weights.  Call sites: tld/diffusion.py:91.
CROSS-CHECKED against an independent third-party implementation of the same network that IS in the image: torchtitan's
FLUX `AutoEncoder` (the CompVis / LDM autoencoder diffusers' AutoencoderKL ports), encoder and decoder, random weights
renamed with diffusers' published LDM key map, fp32, 1e-5 (tests/test_vae_oracle_crosscheck.py).  That verifies the
architecture restatement, not diffusers' code or the sdxl-vae-fp16-fix weights.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def group_norm(x, w, b, groups=32, eps=1e-6):
    B, C, H, W = x.shape
    g = x.reshape(B, groups, (C // groups) * H * W)
    mu = g.mean(dim=2, keepdim=True)
    var = ((g - mu) ** 2).mean(dim=2, keepdim=True)
    g = (g - mu) / torch.sqrt(var + eps)
    return g.reshape(B, C, H, W) * w.view(1, C, 1, 1) + b.view(1, C, 1, 1)


def silu(x):
    return x / (1.0 + torch.exp(-x))


def conv(sd, name, x, pad):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], padding=pad)


def resnet(sd, name, x):
    h = conv(sd, name + ".conv1", silu(group_norm(x, sd[name + ".norm1.weight"], sd[name + ".norm1.bias"])), 1)
    h = conv(sd, name + ".conv2", silu(group_norm(h, sd[name + ".norm2.weight"], sd[name + ".norm2.bias"])), 1)
    if name + ".conv_shortcut.weight" in sd:
        x = conv(sd, name + ".conv_shortcut", x, 0)
    return x + h


def mid_attention(sd, name, x):
    B, C, H, W = x.shape
    h = group_norm(x, sd[name + ".group_norm.weight"], sd[name + ".group_norm.bias"])
    t = h.reshape(B, C, H * W).transpose(1, 2)
    q = t @ sd[name + ".to_q.weight"].t() + sd[name + ".to_q.bias"]
    k = t @ sd[name + ".to_k.weight"].t() + sd[name + ".to_k.bias"]
    v = t @ sd[name + ".to_v.weight"].t() + sd[name + ".to_v.bias"]
    p = torch.softmax((q @ k.transpose(1, 2)) / math.sqrt(C), dim=-1)
    o = (p @ v) @ sd[name + ".to_out.0.weight"].t() + sd[name + ".to_out.0.bias"]
    return x + o.transpose(1, 2).reshape(B, C, H, W)


def upsample_nearest2(x):
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def decode(sd, z, n_up_blocks=4, layers_per_block=2):
    x = conv(sd, "post_quant_conv", z, 0)
    x = conv(sd, "decoder.conv_in", x, 1)
    x = resnet(sd, "decoder.mid_block.resnets.0", x)
    x = mid_attention(sd, "decoder.mid_block.attentions.0", x)
    x = resnet(sd, "decoder.mid_block.resnets.1", x)
    for i in range(n_up_blocks):
        for j in range(layers_per_block + 1):
            x = resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", x)
        if i != n_up_blocks - 1:
            x = conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", upsample_nearest2(x), 1)
    x = silu(group_norm(x, sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"]))
    return conv(sd, "decoder.conv_out", x, 1)


# ---------------------------------------------------------------------------------------------------------------
# Encoder half (call sites tld/data.py:35-41,168).  Same status as the decoder: third-party
# diffusers code, not pinned to diffusers (cross-checked against the LDM AutoEncoder, see the header); restated from the published AutoencoderKL encoder (conv_in 3->128, four down blocks of
# two ResNets with an asymmetric-pad stride-2 conv between them, mid block with one single-head attention,
# GroupNorm/SiLU/conv_out to 2*latent channels, 1x1 quant_conv, diagonal Gaussian with logvar clamped to [-30, 20]).
# ---------------------------------------------------------------------------------------------------------------
def downsample(sd, name, x):
    x = F.pad(x, (0, 1, 0, 1))  # diffusers Downsample2D with padding=0: pad right/bottom by one, then stride-2 conv
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=2)


def encode_moments(sd, x, n_down_blocks=4, layers_per_block=2):
    """image x [B,3,H,W] in [-1,1] -> (mean, logvar) of the latent posterior, each [B,4,H/8,W/8]"""
    h = conv(sd, "encoder.conv_in", x, 1)
    for i in range(n_down_blocks):
        for j in range(layers_per_block):
            h = resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", h)
        if i != n_down_blocks - 1:
            h = downsample(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", h)
    h = resnet(sd, "encoder.mid_block.resnets.0", h)
    h = mid_attention(sd, "encoder.mid_block.attentions.0", h)
    h = resnet(sd, "encoder.mid_block.resnets.1", h)
    h = silu(group_norm(h, sd["encoder.conv_norm_out.weight"], sd["encoder.conv_norm_out.bias"]))
    h = conv(sd, "encoder.conv_out", h, 1)
    moments = conv(sd, "quant_conv", h, 0)
    mean, logvar = moments.chunk(2, dim=1)
    return mean, logvar.clamp(-30.0, 20.0)


def gaussian_sample(mean, logvar, eps):
    """DiagonalGaussianDistribution.sample with the noise injected: mean + exp(0.5 logvar) * eps"""
    return mean + torch.exp(0.5 * logvar) * eps


# uint8 latent storage of the reference's dataset files: restatement of tld/data.py:51-60 (checked against the reference's
# own source in tests/golden/make_golden.py -> tests/golden/latent_quant.npz, so THIS part is pinned)
def quantize_latents(lat, clip_val=20):
    """latent -> byte: clip to [-clip_val, clip_val], map affinely onto [0, 255], truncate (tld/data.py:51-54)"""
    unit = lat.clip(-clip_val, clip_val) / clip_val
    return (255 * ((unit + 1) / 2)).to(torch.uint8)


def dequantize_latents(q, clip_val=20):
    """byte -> fp16 latent, every step rounded to fp16 as torch does on a half tensor (tld/data.py:57-60)"""
    unit = 2 * (q.to(torch.float16) / 255) - 1
    return clip_val * unit
