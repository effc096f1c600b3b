"""``Denoiser`` — drop-in for ``tld.denoiser.Denoiser`` (reference tld/denoiser.py:85-126) whose forward runs in
hand-written sm_100a CUDA behind the C ABI of ``libtld_b200.so``.

Same constructor arguments, attributes (``image_size``, ``noise_embed_dims``, ``embed_dim``, ``n_channels``),
``state_dict`` keys/shapes/order (so the published checkpoint loads with ``strict=True``) and
``forward(x, noise_level, label)`` contract.  The module owns ordinary ``nn.Parameter``s; the library keeps a
packed copy (bf16 GEMM operands, fp32 everything else) that is refreshed whenever a parameter changes.

There is no eager/CPU path here on purpose: CPU tensors raise.  The fp32 restatement used by the tests lives in
``oracle/`` and is never imported by this package.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Iterator, Tuple

import torch
from torch import nn

from . import _lib

_TB = "denoiser_trans_block."


def param_layout(image_size: int, noise_embed_dims: int, patch_size: int, embed_dim: int, n_layers: int,
                 text_emb_size: int, mlp_multiplier: int, n_channels: int) -> "Dict[str, Tuple[tuple, str]]":
    """Ordered {state_dict key: (shape, kind)} of the reference module tree.

    kind: 'linear_w' / 'conv_w' (kaiming-uniform a=sqrt(5)), 'bias:<fan_in>', 'ones', 'zeros', 'normal' (nn.Embedding),
    'speeds' (buffer, transformer_blocks.py:11-15), 'arange' (buffer, denoiser.py:55).
    """
    D, E, Te, p = embed_dim, noise_embed_dims, text_emb_size, patch_size
    pd = n_channels * p * p
    N = int((image_size / p) * (image_size / p))
    H4 = mlp_multiplier * D
    lay: Dict[str, Tuple[tuple, str]] = {}
    lay["fourier_feats.0.angular_speeds"] = ((E // 2,), "speeds")
    lay["fourier_feats.1.weight"] = ((D, E), "linear_w")
    lay["fourier_feats.1.bias"] = ((D,), f"bias:{E}")
    lay["fourier_feats.3.weight"] = ((D, D), "linear_w")
    lay["fourier_feats.3.bias"] = ((D,), f"bias:{D}")
    lay[_TB + "precomputed_pos_enc"] = ((N,), "arange")
    lay[_TB + "patchify_and_embed.0.weight"] = ((pd, n_channels, p, p), "conv_w")
    lay[_TB + "patchify_and_embed.0.bias"] = ((pd,), f"bias:{pd}")
    lay[_TB + "patchify_and_embed.2.weight"] = ((pd,), "ones")
    lay[_TB + "patchify_and_embed.2.bias"] = ((pd,), "zeros")
    lay[_TB + "patchify_and_embed.3.weight"] = ((D, pd), "linear_w")
    lay[_TB + "patchify_and_embed.3.bias"] = ((D,), f"bias:{pd}")
    lay[_TB + "patchify_and_embed.4.weight"] = ((D,), "ones")
    lay[_TB + "patchify_and_embed.4.bias"] = ((D,), "zeros")
    lay[_TB + "pos_embed.weight"] = ((N, D), "normal")
    for i in range(n_layers):
        b = f"{_TB}decoder_blocks.{i}."
        lay[b + "self_attention.qkv_linear.weight"] = ((3 * D, D), "linear_w")
        lay[b + "cross_attention.kv_linear.weight"] = ((2 * D, D), "linear_w")
        lay[b + "cross_attention.q_linear.weight"] = ((D, D), "linear_w")
        lay[b + "mlp.mlp.0.weight"] = ((H4, D, 1, 1), "conv_w")
        lay[b + "mlp.mlp.0.bias"] = ((H4,), f"bias:{D}")
        lay[b + "mlp.mlp.1.weight"] = ((H4, 1, 3, 3), "conv_w")
        lay[b + "mlp.mlp.1.bias"] = ((H4,), "bias:9")
        lay[b + "mlp.mlp.3.weight"] = ((D, H4, 1, 1), "conv_w")
        lay[b + "mlp.mlp.3.bias"] = ((D,), f"bias:{H4}")
        for n in ("norm1", "norm2", "norm3"):
            lay[b + n + ".weight"] = ((D,), "ones")
            lay[b + n + ".bias"] = ((D,), "zeros")
    lay[_TB + "out_proj.0.weight"] = ((pd, D), "linear_w")
    lay[_TB + "out_proj.0.bias"] = ((pd,), f"bias:{D}")
    lay["norm.weight"] = ((D,), "ones")
    lay["norm.bias"] = ((D,), "zeros")
    lay["label_proj.weight"] = ((D, Te), "linear_w")
    lay["label_proj.bias"] = ((D,), f"bias:{Te}")
    return lay


def _init_tensor(shape, kind: str) -> torch.Tensor:
    """Default torch initialisers of the reference's layers (nn.Linear/Conv2d/LayerNorm/Embedding)."""
    if kind in ("linear_w", "conv_w"):
        fan_in = int(math.prod(shape[1:]))
        bound = 1.0 / math.sqrt(fan_in)  # kaiming_uniform_(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
        return torch.empty(shape).uniform_(-bound, bound)
    if kind.startswith("bias:"):
        bound = 1.0 / math.sqrt(int(kind.split(":")[1]))
        return torch.empty(shape).uniform_(-bound, bound)
    if kind == "ones":
        return torch.ones(shape)
    if kind == "zeros":
        return torch.zeros(shape)
    if kind == "normal":
        return torch.randn(shape)
    if kind == "speeds":
        n = shape[0]
        freqs = torch.exp(torch.linspace(math.log(1.0), math.log(1000.0), n))
        return 2.0 * torch.pi * freqs
    if kind == "arange":
        return torch.arange(0, shape[0]).long()
    raise ValueError(kind)


class _Node(nn.Module):
    """Anonymous container; the tree of these reproduces the reference's dotted state_dict names."""


def _attach(root: nn.Module, dotted: str, value: torch.Tensor, buffer: bool) -> None:
    *path, leaf = dotted.split(".")
    mod = root
    for name in path:
        if name not in mod._modules:
            mod.add_module(name, _Node())
        mod = mod._modules[name]
    if buffer:
        mod.register_buffer(leaf, value)
    else:
        mod.register_parameter(leaf, nn.Parameter(value))


class Denoiser(nn.Module):
    """Transformer denoiser: patchify -> n_layers x (self-attn, 2-token cross-attn, sep-conv MLP) -> unpatchify."""

    def __init__(self, image_size: int, noise_embed_dims: int, patch_size: int, embed_dim: int, dropout: float,
                 n_layers: int, text_emb_size: int = 768, mlp_multiplier: int = 4, n_channels: int = 4):
        super().__init__()
        self.image_size = image_size
        self.noise_embed_dims = noise_embed_dims
        self.embed_dim = embed_dim
        self.n_channels = n_channels
        self.patch_size = patch_size
        self.dropout = dropout
        self.n_layers = n_layers
        self.text_emb_size = text_emb_size
        self.mlp_multiplier = mlp_multiplier
        self._layout = param_layout(image_size, noise_embed_dims, patch_size, embed_dim, n_layers, text_emb_size,
                                    mlp_multiplier, n_channels)
        for key, (shape, kind) in self._layout.items():
            _attach(self, key, _init_tensor(shape, kind), buffer=kind in ("speeds", "arange"))
        # library state (never part of state_dict / deepcopy)
        self.__dict__["_handle"] = None
        self.__dict__["_handle_device"] = None
        self.__dict__["_checked_complete"] = False

    # ------------------------------------------------------------------ library handle management
    def __deepcopy__(self, memo):
        clone = Denoiser(self.image_size, self.noise_embed_dims, self.patch_size, self.embed_dim, self.dropout,
                         self.n_layers, self.text_emb_size, self.mlp_multiplier, self.n_channels)
        ref = next(self.parameters())
        clone.to(device=ref.device, dtype=ref.dtype)
        clone.load_state_dict(self.state_dict())
        clone.train(self.training)
        for pc, ps in zip(clone.parameters(), self.parameters()):
            pc.requires_grad_(ps.requires_grad)
        return clone

    def __del__(self):
        self._release()

    def _release(self):
        h = self.__dict__.get("_handle")
        if h:
            try:
                _lib.load().tld_denoiser_destroy(h)
            except Exception:
                pass
            self.__dict__["_handle"] = None

    def _float_entries(self) -> "Iterator[Tuple[str, torch.Tensor]]":
        for key, t in self.state_dict(keep_vars=True).items():
            if t.is_floating_point():
                yield key, t

    def _ensure_handle(self, device: torch.device):
        lib = _lib.load()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if self.__dict__["_handle"] is None or self.__dict__["_handle_device"] != idx:
            self._release()
            cfg = _lib.TldConfig(self.image_size, self.noise_embed_dims, self.patch_size, self.embed_dim,
                                 self.n_layers, self.text_emb_size, self.mlp_multiplier, self.n_channels, 0.0)
            h = C.c_void_p()
            _lib.check(lib.tld_denoiser_create(C.byref(cfg), idx, C.byref(h)), "tld_denoiser_create")
            self.__dict__["_handle"] = h
            self.__dict__["_handle_device"] = idx
            self.__dict__["_checked_complete"] = False
        # Refresh the packed (bf16 / transposed) weights from the fp32 parameters on EVERY call, asynchronously on the
        # current stream.  A cached copy cannot be trusted: in-place updates are not reliably visible in torch's version
        # counters (torch.optim.Adam(fused=True) leaves them untouched, and so does the reference's own
        # `ema_param.data.mul_(alpha).add_(...)`, tld/train.py:55-58).  One pass over the weights, ~0.15 ms on a B200.
        h = self.__dict__["_handle"]
        dev = torch.device("cuda", idx)
        entries = list(self._float_entries())
        keys = tuple(k for k, _ in entries)
        cache = self.__dict__.get("_key_arrays")
        if cache is None or cache[0] != keys:
            arr = (C.c_char_p * len(keys))(*[k.encode() for k in keys])
            cache = (keys, arr, (C.c_int64 * len(keys))(*[t.numel() for _, t in entries]))
            self.__dict__["_key_arrays"] = cache
        srcs = []
        for _, t in entries:
            s = t.detach()
            if s.device != dev or s.dtype != torch.float32 or not s.is_contiguous():
                s = s.to(device=dev, dtype=torch.float32).contiguous()   # temporary, ordered on the same stream
            srcs.append(s)
        ptrs = (C.c_void_p * len(srcs))(*[s.data_ptr() for s in srcs])
        with torch.cuda.device(dev):
            _lib.check(lib.tld_denoiser_set_params_async(h, len(srcs), cache[1], ptrs, cache[2], _lib.current_stream_ptr(dev)),
                       "tld_denoiser_set_params_async")
        if not self.__dict__.get("_checked_complete"):
            missing = lib.tld_denoiser_missing_params(h)
            if missing:
                raise _lib.TldError(f"{missing} parameters were not provided to the library")
            self.__dict__["_checked_complete"] = True
        return self.__dict__["_handle"]

    # ------------------------------------------------------------------ forward
    def check_inputs(self, x: torch.Tensor, noise_level: torch.Tensor, label: torch.Tensor) -> int:
        """Shape / device validation shared by the inference and the training forward (the library reads raw pointers:
        a wrong label width would be an out-of-bounds read). Returns the batch size."""
        if not x.is_cuda:
            raise _lib.TldError("transformer_latent_diffusion_b200.Denoiser runs on CUDA (sm_100a) only; "
                                "there is no CPU fallback")
        B = x.shape[0]
        if x.dim() != 4 or tuple(x.shape[1:]) != (self.n_channels, self.image_size, self.image_size):
            raise ValueError(f"expected x of shape [B,{self.n_channels},{self.image_size},{self.image_size}], "
                             f"got {tuple(x.shape)}")
        if noise_level.numel() != B or label.dim() != 2 or label.shape[0] != B or label.shape[-1] != self.text_emb_size:
            raise ValueError("noise_level must be [B,1] and label [B,text_emb_size]")
        if self.dropout != 0 and self.training:
            # transformer_blocks.py:43,105 apply dropout only in training mode; eval-mode sampling with dropout > 0 in the
            # config is fine, training with it is not implemented (checked on every call: .train() can come later)
            raise _lib.TldError("dropout > 0 in training mode is not supported by the B200 path")
        return B

    def forward(self, x: torch.Tensor, noise_level: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
        """x[B,C,H,W], noise_level[B,1], label[B,text_emb] -> x0 prediction [B,C,H,W] (tld/denoiser.py:116-126)."""
        B = self.check_inputs(x, noise_level, label)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .train import denoiser_autograd_forward  # backward kernels live with the train step

            return denoiser_autograd_forward(self, x, noise_level, label)
        h = self._ensure_handle(x.device)
        xf = x.detach().to(torch.float32).contiguous()
        tf = noise_level.detach().to(device=x.device, dtype=torch.float32).reshape(B).contiguous()
        lf = label.detach().to(device=x.device, dtype=torch.float32).contiguous()
        out = torch.empty_like(xf)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().tld_denoiser_forward(h, _lib.ptr(xf), _lib.ptr(tf), _lib.ptr(lf), _lib.ptr(out), B,
                                                        _lib.current_stream_ptr(x.device)), "tld_denoiser_forward")
        return out.to(x.dtype)
