"""The four tensor helpers of ``tld/data.py`` that sit directly on either side of the hot path:
``encode_image`` / ``decode_latents`` (tld/data.py:34-48) and the uint8 latent ``quantize_latents`` /
``dequantize_latents`` (tld/data.py:51-60) that the reference's dataset files are stored with.  The rest of ``tld/data.py``
(img2dataset download, webdataset/HDF5 plumbing, CLIP text encoding) is out of scope.

B200-native differences, same results: the (de)quantisers run as CUDA kernels on device tensors
(``tld_latent_quantize`` / ``tld_latent_dequantize``, bit-exact with the reference's arithmetic - golden vectors from the
reference's own source in tests/golden/latent_quant.npz), so a latent leaves the GPU as 1 byte per value instead of the
fp16 the reference copies to the host first.  CPU tensors raise: there is no CPU fallback.
"""
from __future__ import annotations

import torch
from torch import Tensor

from . import _lib


def _need_cuda(t: Tensor, what: str) -> None:
    if not t.is_cuda:
        raise _lib.TldError(f"{what}: CUDA tensor required (no CPU fallback)")


@torch.no_grad()
def encode_image(img: Tensor, vae, generator=None, to_cpu: bool = True) -> Tensor:
    """img [B,3,H,W] in [0,1] -> sampled latent [B,4,H/8,W/8] (unscaled), fp16, as tld/data.py:34-41.
    ``to_cpu=False`` keeps the latent on the device (e.g. to quantise it there)."""
    dev = next(vae.parameters()).device
    x = img.to(dev).to(torch.float16)
    x = x * 2 - 1  # to [-1, 1] (tld/data.py:38)
    encoded = vae.encode(x, return_dict=False)[0].sample(generator)
    return encoded.cpu() if to_cpu else encoded


@torch.no_grad()
def decode_latents(out_latents: Tensor, vae) -> Tensor:
    """unscaled latents -> images in [0,1] on the CPU (tld/data.py:44-48)"""
    dev = next(vae.parameters()).device
    out = vae.decode(out_latents.to(dev))[0].cpu()
    return ((out + 1) / 2).clip(0, 1)


def quantize_latents(lat: Tensor, clip_val: float = 20) -> Tensor:
    """scale and quantise latents to uint8 (tld/data.py:51-54): clip to [-clip_val, clip_val], map to [0, 255], truncate."""
    _need_cuda(lat, "quantize_latents")
    if lat.dtype not in (torch.float32, torch.float16):
        lat = lat.float()
    lat = lat.contiguous()
    out = torch.empty(lat.shape, dtype=torch.uint8, device=lat.device)
    _lib.check(_lib.load().tld_latent_quantize(_lib.ptr(lat), int(lat.dtype == torch.float16), _lib.ptr(out), lat.numel(),
                                               float(clip_val), _lib.current_stream_ptr(lat.device)), "tld_latent_quantize")
    return out


def dequantize_latents(lat: Tensor, clip_val: float = 20) -> Tensor:
    """uint8 -> fp16 latents (tld/data.py:57-60)"""
    _need_cuda(lat, "dequantize_latents")
    if lat.dtype != torch.uint8:
        raise ValueError("dequantize_latents: uint8 tensor expected")
    lat = lat.contiguous()
    out = torch.empty(lat.shape, dtype=torch.float16, device=lat.device)
    _lib.check(_lib.load().tld_latent_dequantize(_lib.ptr(lat), _lib.ptr(out), lat.numel(), float(clip_val),
                                                 _lib.current_stream_ptr(lat.device)), "tld_latent_dequantize")
    return out
