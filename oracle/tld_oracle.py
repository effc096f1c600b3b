"""CPU oracle for the transformer-latent-diffusion hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (plain torch CPU ops over a flat ``state_dict``) of the
reference algorithm, written so that every function can be checked line by line against
``/root/reference/tld``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The shipped package never does: the
product path is the CUDA library behind ``include/tld_b200.h`` and fails loudly without it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified reference from
``/root/reference`` (in the build container), runs it on seeded inputs and commits the outputs
under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them through this file.
The one un-pinned piece is the VAE decoder (``oracle/vae_oracle.py``): its reference is the
third-party ``diffusers`` package, which is absent here.

Every function cites the reference ``file:line`` it follows.  Arithmetic dtype is whatever the
inputs are (fp32 in the tests, fp64 for noise-floor measurements).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

HEAD_DIM = 64  # reference: n_heads = embed_dim // 64 (tld/transformer_blocks.py:126-129)
LN_EPS = 1e-5  # nn.LayerNorm default, used everywhere in the reference


@dataclass(frozen=True)
class OracleCfg:
    """Mirror of the constructor arguments of ``Denoiser`` (tld/denoiser.py:86-97)."""

    image_size: int = 16
    noise_embed_dims: int = 256
    patch_size: int = 2
    embed_dim: int = 128
    dropout: float = 0.0
    n_layers: int = 3
    text_emb_size: int = 768
    mlp_multiplier: int = 4
    n_channels: int = 4

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def seq_len(self) -> int:  # tld/denoiser.py:31
        return int((self.image_size / self.patch_size) * (self.image_size / self.patch_size))

    @property
    def patch_dim(self) -> int:  # tld/denoiser.py:32
        return self.n_channels * self.patch_size * self.patch_size

    @property
    def heads(self) -> int:
        return self.embed_dim // HEAD_DIM


# --------------------------------------------------------------------------------------------
# parameter inventory + deterministic synthetic weights
# --------------------------------------------------------------------------------------------
_TB = "denoiser_trans_block."


def param_shapes(cfg: OracleCfg) -> "dict[str, tuple]":
    """state_dict keys and shapes of the reference ``Denoiser`` (SURVEY.md §8a 'Weights')."""
    D, pd, N, H4 = cfg.embed_dim, cfg.patch_dim, cfg.seq_len, cfg.mlp_multiplier * cfg.embed_dim
    p = cfg.patch_size
    out: dict[str, tuple] = {
        "fourier_feats.0.angular_speeds": (cfg.noise_embed_dims // 2,),
        "fourier_feats.1.weight": (D, cfg.noise_embed_dims),
        "fourier_feats.1.bias": (D,),
        "fourier_feats.3.weight": (D, D),
        "fourier_feats.3.bias": (D,),
        _TB + "precomputed_pos_enc": (N,),
        _TB + "patchify_and_embed.0.weight": (pd, cfg.n_channels, p, p),
        _TB + "patchify_and_embed.0.bias": (pd,),
        _TB + "patchify_and_embed.2.weight": (pd,),
        _TB + "patchify_and_embed.2.bias": (pd,),
        _TB + "patchify_and_embed.3.weight": (D, pd),
        _TB + "patchify_and_embed.3.bias": (D,),
        _TB + "patchify_and_embed.4.weight": (D,),
        _TB + "patchify_and_embed.4.bias": (D,),
        _TB + "pos_embed.weight": (N, D),
    }
    for i in range(cfg.n_layers):
        b = f"{_TB}decoder_blocks.{i}."
        out[b + "self_attention.qkv_linear.weight"] = (3 * D, D)
        out[b + "cross_attention.kv_linear.weight"] = (2 * D, D)
        out[b + "cross_attention.q_linear.weight"] = (D, D)
        out[b + "mlp.mlp.0.weight"] = (H4, D, 1, 1)
        out[b + "mlp.mlp.0.bias"] = (H4,)
        out[b + "mlp.mlp.1.weight"] = (H4, 1, 3, 3)
        out[b + "mlp.mlp.1.bias"] = (H4,)
        out[b + "mlp.mlp.3.weight"] = (D, H4, 1, 1)
        out[b + "mlp.mlp.3.bias"] = (D,)
        for n in ("norm1", "norm2", "norm3"):
            out[b + n + ".weight"] = (D,)
            out[b + n + ".bias"] = (D,)
    out[_TB + "out_proj.0.weight"] = (pd, D)
    out[_TB + "out_proj.0.bias"] = (pd,)
    out["norm.weight"] = (D,)
    out["norm.bias"] = (D,)
    out["label_proj.weight"] = (D, cfg.text_emb_size)
    out["label_proj.bias"] = (D,)
    return out


def angular_speeds(noise_embed_dims: int) -> torch.Tensor:
    """tld/transformer_blocks.py:11-15 — 2*pi*exp(linspace(ln 1, ln 1000, E/2)), fp32."""
    freqs = torch.exp(torch.linspace(np.log(1.0), np.log(1000.0), noise_embed_dims // 2))
    return 2.0 * torch.pi * freqs


def synth_state_dict(cfg: OracleCfg, seed: int, dtype=torch.float32) -> "dict[str, torch.Tensor]":
    """Deterministic synthetic weights shared by golden generation and every parity test.

    Not the reference's init: every matrix is N(0, 1/fan_in)-ish so activations stay O(1) through
    12 blocks, LayerNorm affine parameters are perturbed away from (1, 0) so they are exercised,
    and all biases are non-zero.  The stream is torch's CPU generator, which is identical in this
    container and on the GPU box (same image).
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd: dict[str, torch.Tensor] = {}
    for key, shape in param_shapes(cfg).items():
        if key.endswith("angular_speeds"):
            sd[key] = angular_speeds(cfg.noise_embed_dims)
        elif key.endswith("precomputed_pos_enc"):
            sd[key] = torch.arange(0, cfg.seq_len).long()
        elif ("norm" in key or key.endswith("patchify_and_embed.2.weight")
              or key.endswith("patchify_and_embed.4.weight")
              or key.endswith("patchify_and_embed.2.bias")
              or key.endswith("patchify_and_embed.4.bias")) and len(shape) == 1:
            r = torch.randn(shape, generator=g)
            sd[key] = (1.0 + 0.1 * r) if key.endswith("weight") else 0.1 * r
        elif key.endswith("bias"):
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("pos_embed.weight"):
            sd[key] = torch.randn(shape, generator=g)
        elif key.endswith("mlp.mlp.1.weight"):
            sd[key] = torch.randn(shape, generator=g) / 3.0
        else:
            fan_in = int(np.prod(shape[1:]))
            sd[key] = torch.randn(shape, generator=g) / math.sqrt(fan_in)
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


# --------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------
def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """nn.LayerNorm over the last dim, biased variance, eps 1e-5 (transformer_blocks.py:131-133)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + LN_EPS) * w + b


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    """nn.GELU() default = exact erf form (transformer_blocks.py:103, denoiser.py:108)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def sinusoidal_embedding(speeds: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """tld/transformer_blocks.py:17-21 — cat[sin(w t), cos(w t)] along the last dim; t is [B,1]."""
    arg = speeds * t
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)


def cond_tokens(sd, cfg: OracleCfg, noise_level: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """tld/denoiser.py:117-122 — the 2-token conditioning sequence y[B,2,D] (noise tok, label tok)."""
    e = sinusoidal_embedding(sd["fourier_feats.0.angular_speeds"], noise_level)
    n = F.linear(e, sd["fourier_feats.1.weight"], sd["fourier_feats.1.bias"])
    n = gelu_erf(n)
    n = F.linear(n, sd["fourier_feats.3.weight"], sd["fourier_feats.3.bias"])
    lab = F.linear(label, sd["label_proj.weight"], sd["label_proj.bias"])
    y = torch.stack([n, lab], dim=1)
    return layer_norm(y, sd["norm.weight"], sd["norm.bias"])


def patch_embed(sd, cfg: OracleCfg, x: torch.Tensor) -> torch.Tensor:
    """tld/denoiser.py:34-45,75-77 — strided conv -> tokens -> LN(pd) -> Linear -> LN(D) -> +pos."""
    B, C, Hh, Ww = x.shape
    p, g = cfg.patch_size, cfg.grid
    # conv k=s=p is a per-patch matvec: gather [B, g, g, C, p, p] patches
    patches = x.reshape(B, C, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, C * p * p)
    w0 = sd[_TB + "patchify_and_embed.0.weight"].reshape(cfg.patch_dim, -1)
    t = patches @ w0.t() + sd[_TB + "patchify_and_embed.0.bias"]
    t = layer_norm(t, sd[_TB + "patchify_and_embed.2.weight"], sd[_TB + "patchify_and_embed.2.bias"])
    t = F.linear(t, sd[_TB + "patchify_and_embed.3.weight"], sd[_TB + "patchify_and_embed.3.bias"])
    t = layer_norm(t, sd[_TB + "patchify_and_embed.4.weight"], sd[_TB + "patchify_and_embed.4.bias"])
    return t + sd[_TB + "pos_embed.weight"][: t.shape[1]]


def _split_heads(t: torch.Tensor, heads: int) -> torch.Tensor:
    B, n, _ = t.shape
    return t.reshape(B, n, heads, HEAD_DIM).permute(0, 2, 1, 3)  # "bs n (h d) -> bs h n d"


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """tld/transformer_blocks.py:31-48 — non-causal softmax(q k^T / sqrt(64)) v, heads merged back."""
    qh, kh, vh = (_split_heads(t, heads) for t in (q, k, v))
    outs = []
    for b in range(qh.shape[0]):  # per sample keeps the N=4096 score matrix at heads*N*N
        s = (qh[b] @ kh[b].transpose(-1, -2)) * (1.0 / math.sqrt(HEAD_DIM))
        outs.append(torch.softmax(s, dim=-1) @ vh[b])
    o = torch.stack(outs)
    B, h, n, d = o.shape
    return o.permute(0, 2, 1, 3).reshape(B, n, h * d)


def self_attention(x: torch.Tensor, wqkv: torch.Tensor, heads: int) -> torch.Tensor:
    """tld/transformer_blocks.py:57-59 — bias-free qkv linear, chunk(3), attention, NO out-proj."""
    q, k, v = F.linear(x, wqkv).chunk(3, dim=2)
    return attention(q, k, v, heads)


def cross_attention(x, y, wq, wkv, heads: int) -> torch.Tensor:
    """tld/transformer_blocks.py:69-72 — q from x, k/v from the 2 cond tokens, NO out-proj."""
    q = F.linear(x, wq)
    k, v = F.linear(y, wkv).chunk(2, dim=2)
    return attention(q, k, v, heads)


def mlp_sepconv(x: torch.Tensor, w0, b0, w1, b1, w3, b3) -> torch.Tensor:
    """tld/transformer_blocks.py:108-113 — 1x1 conv, depthwise 3x3 'same' (zero pad), GELU, 1x1 conv."""
    B, n, D = x.shape
    g = int(np.sqrt(n))
    hid = F.linear(x, w0.reshape(w0.shape[0], D), b0)  # 1x1 conv == per-token linear
    grid = hid.reshape(B, g, g, -1).permute(0, 3, 1, 2)  # "bs (h w) d -> bs d h w"
    grid = F.conv2d(grid, w1, b1, padding=1, groups=w1.shape[0])
    hid = gelu_erf(grid).permute(0, 2, 3, 1).reshape(B, n, -1)
    return F.linear(hid, w3.reshape(w3.shape[0], -1), b3)


def decoder_block(sd, cfg: OracleCfg, i: int, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """tld/transformer_blocks.py:135-139 — three pre-LN residual sub-layers."""
    b = f"{_TB}decoder_blocks.{i}."
    x = self_attention(layer_norm(x, sd[b + "norm1.weight"], sd[b + "norm1.bias"]),
                       sd[b + "self_attention.qkv_linear.weight"], cfg.heads) + x
    x = cross_attention(layer_norm(x, sd[b + "norm2.weight"], sd[b + "norm2.bias"]), y,
                        sd[b + "cross_attention.q_linear.weight"],
                        sd[b + "cross_attention.kv_linear.weight"], cfg.heads) + x
    x = mlp_sepconv(layer_norm(x, sd[b + "norm3.weight"], sd[b + "norm3.bias"]),
                    sd[b + "mlp.mlp.0.weight"], sd[b + "mlp.mlp.0.bias"],
                    sd[b + "mlp.mlp.1.weight"], sd[b + "mlp.mlp.1.bias"],
                    sd[b + "mlp.mlp.3.weight"], sd[b + "mlp.mlp.3.bias"]) + x
    return x


def out_proj(sd, cfg: OracleCfg, x: torch.Tensor) -> torch.Tensor:
    """tld/denoiser.py:47-52,72,82 — Linear D->pd then 'b (h w) (c p1 p2) -> b c (h p1) (w p2)'."""
    t = F.linear(x, sd[_TB + "out_proj.0.weight"], sd[_TB + "out_proj.0.bias"])
    B = t.shape[0]
    g, p, C = cfg.grid, cfg.patch_size, cfg.n_channels
    return t.reshape(B, g, g, C, p, p).permute(0, 3, 1, 4, 2, 5).reshape(B, C, g * p, g * p)


def denoiser_forward(sd, cfg: OracleCfg, x, noise_level, label, return_intermediates=False):
    """tld/denoiser.py:116-126 (+ :74-82) — the full Denoiser.forward."""
    y = cond_tokens(sd, cfg, noise_level, label)
    h = patch_embed(sd, cfg, x)
    inter = {"cond": y, "embed": h}
    for i in range(cfg.n_layers):
        h = decoder_block(sd, cfg, i, h, y)
        inter[f"block{i}"] = h
    out = out_proj(sd, cfg, h)
    return (out, inter) if return_intermediates else out


# --------------------------------------------------------------------------------------------
# sampler (tld/diffusion.py:29-125)
# --------------------------------------------------------------------------------------------
def noise_schedule(n_iter: int, exponent: float = 1, noise_levels=None) -> "list[float]":
    """tld/diffusion.py:50-52 — fp32 arange/pow, converted to python floats; [0] forced to 0.99."""
    if noise_levels is None:
        noise_levels = (1 - torch.pow(torch.arange(0, 1, 1 / n_iter), exponent)).tolist()
    noise_levels = list(noise_levels)
    noise_levels[0] = 0.99
    return noise_levels


def multistep_ratios(noise_levels) -> "list[float]":
    """tld/diffusion.py:54-57 — log-SNR steps h_i and their ratios r_i (float64 on the host)."""
    lambdas = [np.log((1 - s) / s) for s in noise_levels]
    hs = [lambdas[i] - lambdas[i - 1] for i in range(1, len(lambdas))]
    return [hs[i - 1] / hs[i] for i in range(1, len(hs))]


def cfg_combine(x0_both: torch.Tensor, num_imgs: int, g: float) -> torch.Tensor:
    """tld/diffusion.py:122-125 — cond half first, uncond half second."""
    return g * x0_both[:num_imgs] + (1 - g) * x0_both[num_imgs:]


def generate_latents(sd, cfg: OracleCfg, labels: torch.Tensor, seeds: torch.Tensor, n_iter: int = 30,
                     class_guidance: float = 3, exponent: float = 1, sharp_f: float = 0.1,
                     bright_f: float = 0.1, noise_levels=None, use_ddpm_plus: bool = True,
                     model=None, trace=None) -> torch.Tensor:
    """tld/diffusion.py:29-89 — reverse sampling up to (not including) the VAE decode.

    ``seeds`` is the initial noise (the reference's ``seeds=`` argument, diffusion.py:119-120).
    ``model`` overrides the denoiser call (used to time the loop around another implementation).
    ``trace`` (a list) collects (x_t, sigma, x0_pred) per model call for teacher-forced parity.
    """
    sig = noise_schedule(n_iter, exponent, noise_levels)
    rs = multistep_ratios(sig) if use_ddpm_plus else None
    x_t = seeds.clone()
    num = x_t.shape[0]
    lab2 = torch.cat([labels, torch.zeros_like(labels)])  # diffusion.py:61
    f = model if model is not None else (lambda x, t, l: denoiser_forward(sd, cfg, x, t, l))

    def pred(xt, s):  # diffusion.py:94-103
        t = torch.full((2 * num, 1), s, dtype=xt.dtype, device=xt.device)
        x0 = cfg_combine(f(torch.cat([xt, xt]), t, lab2), num, class_guidance)
        if trace is not None:
            trace.append((xt.clone(), s, x0.clone()))
        return x0

    prev = None
    nxt = sig[0]
    for i in range(len(sig) - 1):  # diffusion.py:66-83
        cur, nxt = sig[i], sig[i + 1]
        x0 = pred(x_t, cur)
        if prev is None or not use_ddpm_plus:
            d = x0
        else:
            d = (1 + 1 / (2 * rs[i - 1])) * x0 - (1 / (2 * rs[i - 1])) * prev
        x_t = ((cur - nxt) * d + nxt * x_t) / cur
        prev = x0
    x0 = pred(x_t, nxt)  # diffusion.py:85
    x0[:, 3, :, :] += sharp_f  # diffusion.py:88-89
    x0[:, 0, :, :] += bright_f
    return x0


# --------------------------------------------------------------------------------------------
# train step glue (tld/train.py:121-138,160-173,55-58) around an autograd-capable forward
# --------------------------------------------------------------------------------------------
def noise_inputs(x, sigma, eps, label, drop_mask, vae_scale: float = 8.0):
    """tld/train.py:122-138 with sigma, eps and the 15 % label-dropout mask injected as tensors."""
    x = x / vae_scale
    s = sigma.double().view(-1, 1, 1, 1)
    x_noisy = (s * eps.double() + (1 - s) * x.double()).float()
    label = label.clone()
    label[drop_mask] = 0
    return x, x_noisy, sigma.float().view(-1, 1), label


def ema_update(ema: "dict[str, torch.Tensor]", cur: "dict[str, torch.Tensor]", alpha: float = 0.999):
    """tld/train.py:55-58 — ema = alpha*ema + (1-alpha)*param, parameters only (not buffers)."""
    for k in ema:
        if ema[k].is_floating_point() and not k.endswith("angular_speeds"):
            ema[k].mul_(alpha).add_(cur[k], alpha=1 - alpha)


def flops_forward(cfg: OracleCfg) -> float:
    """Algorithmic FLOPs of one sample's forward (SURVEY.md §8d / BASELINE.md §3)."""
    N, D, L = cfg.seq_len, cfg.embed_dim, cfg.n_layers
    pd, E, Te = cfg.patch_dim, cfg.noise_embed_dims, cfg.text_emb_size
    blk = 24 * N * D * D + 4 * N * N * D + 80 * N * D + 8 * D * D
    eh = 2 * N * pd * pd + 4 * N * pd * D + 2 * (E * D + D * D) + 2 * Te * D
    return float(L * blk + eh)
