"""VAE decoder (row A13 of SURVEY.md §8a): the SDXL ``AutoencoderKL`` decoder the reference calls once per
generation (``self.vae.decode(z)[0]``, tld/diffusion.py:91; ``madebyollin/sdxl-vae-fp16-fix``, tld/configs.py:42).

The reference takes this module from the third-party ``diffusers`` package, which (like its weights) is not
available offline, so PARITY IS UNPINNED for this row: the architecture below follows the published
diffusers/SDXL-VAE structure with diffusers-compatible ``state_dict`` keys (``post_quant_conv``,
``decoder.conv_in``, ``decoder.mid_block.*``, ``decoder.up_blocks.*``, ``decoder.conv_norm_out``,
``decoder.conv_out``) so the real checkpoint loads when it is available, and is checked against the independent
fp32 restatement in ``oracle/vae_oracle.py`` with random weights.

STATUS: on CUDA/bf16 every layer runs on libtld_b200 kernels.  GroupNorm(+SiLU) and nearest-2x upsample are hand-written
row-wise kernels (``csrc/vae_kernels.cu``); every 3x3 convolution is an implicit GEMM on the tcgen05 GEMM core
(``tld_vae_conv3x3_fused``: 4-D TMA boxes shifted per tap, zero fill = padding; the ResnetBlock shortcut add and the GroupNorm
statistics partials of its output ride in the epilogue) - the thin ones (conv_in 4 -> 512, conv_out
128 -> 3, the encoder's 3 -> 128 / 512 -> 8) with their channel count zero-padded to 64; the 1x1 convolutions (shortcuts,
post_quant_conv / quant_conv) and the q/k/v/out projections of the mid-block attention are plain tcgen05 GEMMs (``tld_op_gemm``
with the bias epilogue); the encoder's stride-2 convolutions are the stride-1 kernel followed by a 2x sub-sampling; the
softmax(QK^T)V core of the single 512-wide head of the mid-block attention is GEMM -> row softmax -> GEMM per image
(``tld_vae_attention_core``).  There is NO silent fallback: a CUDA tensor the kernels do not cover (fp32, odd channel counts, tiny maps) raises
``TldError``; the plain ATen graph (what the CPU wiring tests and the fp32 oracle cross-check use) only runs when the module
was built with ``allow_aten=True``.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F
from torch import nn

GN_GROUPS = 32
GN_EPS = 1e-6
BLOCK_OUT = (128, 256, 512, 512)  # sdxl-vae config.json block_out_channels
LAYERS_PER_BLOCK = 2
LATENT_CH = 4


def vae_param_layout(latent_ch: int = LATENT_CH, block_out=BLOCK_OUT) -> "Dict[str, tuple]":
    """Ordered {key: shape} of the decoder half of diffusers.AutoencoderKL (decoder + post_quant_conv)."""
    lay: Dict[str, tuple] = {}

    def conv(name, cout, cin, k):
        lay[name + ".weight"] = (cout, cin, k, k)
        lay[name + ".bias"] = (cout,)

    def norm(name, c):
        lay[name + ".weight"] = (c,)
        lay[name + ".bias"] = (c,)

    def resnet(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cout, cin, 3)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".conv_shortcut", cout, cin, 1)

    top = block_out[-1]
    conv("post_quant_conv", latent_ch, latent_ch, 1)
    conv("decoder.conv_in", top, latent_ch, 3)
    resnet("decoder.mid_block.resnets.0", top, top)
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", top)
    for p in ("to_q", "to_k", "to_v", "to_out.0"):
        lay[f"{a}.{p}.weight"] = (top, top)
        lay[f"{a}.{p}.bias"] = (top,)
    resnet("decoder.mid_block.resnets.1", top, top)
    rev = list(reversed(block_out))
    prev = rev[0]
    for i, cout in enumerate(rev):
        for j in range(LAYERS_PER_BLOCK + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else cout, cout)
        if i != len(rev) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
        prev = cout
    norm("decoder.conv_norm_out", block_out[0])
    conv("decoder.conv_out", 3, block_out[0], 3)
    return lay


class _Node(nn.Module):
    pass


def _attach(root: nn.Module, dotted: str, value: torch.Tensor) -> None:
    *path, leaf = dotted.split(".")
    mod = root
    for name in path:
        if name not in mod._modules:
            mod.add_module(name, _Node())
        mod = mod._modules[name]
    mod.register_parameter(leaf, nn.Parameter(value))


class AutoencoderKLDecoder(nn.Module):
    """``decode(z) -> (image,)`` with the SDXL-VAE decoder topology; z is the latent as the reference passes it
    (already multiplied by ``scale_factor``, tld/diffusion.py:91)."""

    def __init__(self, latent_ch: int = LATENT_CH, block_out: Tuple[int, ...] = BLOCK_OUT, chunk: int = 32,
                 allow_aten: bool = False):
        super().__init__()
        self.latent_ch, self.block_out, self.chunk = latent_ch, tuple(block_out), chunk
        self.allow_aten = allow_aten  # test-only: run layers without a libtld_b200 kernel on the plain ATen graph
        self.force_aten = False       # test-only (needs allow_aten): skip the kernels everywhere = torch's own bf16 graph
        self.own_launches = 0  # libtld_b200 kernels launched by decode() so far (bench.py's gpu_launches)
        self.fuse_gn_stats = True  # GroupNorm statistics partials + shortcut add in the producing conv's epilogue
        self._layout = vae_param_layout(latent_ch, block_out)
        for key, shape in self._layout.items():
            if key.endswith("weight") and len(shape) == 1:
                t = torch.ones(shape)
            elif key.endswith("bias"):
                t = torch.zeros(shape)
            else:
                fan_in = int(math.prod(shape[1:]))
                t = torch.empty(shape).uniform_(-1.0 / math.sqrt(fan_in), 1.0 / math.sqrt(fan_in))
            _attach(self, key, t)

    def _p(self, key: str) -> torch.Tensor:
        mod = self
        *path, leaf = key.split(".")
        for name in path:
            mod = mod._modules[name]
        return mod._parameters[leaf]

    # -- building blocks ---------------------------------------------------------------------------------------
    def _aten(self, what: str, x) -> None:
        """gate of every ATen fallback: the product path (CUDA, bf16) never dispatches to library kernels silently"""
        if not self.allow_aten:
            from . import _lib

            raise _lib.TldError(
                f"vae: {what} has no libtld_b200 kernel for a {tuple(x.shape)} {x.dtype} tensor on {x.device} (the kernels take "
                "bf16 CUDA tensors, channel counts that are multiples of 32 with channels/32 a multiple of 4, maps of at least 128 "
                "pixels); the ATen reference graph only runs on a module built with allow_aten=True (tests)")

    def _compute_dtype(self, w) -> torch.dtype:
        """On CUDA the module computes with bf16 tensor-core operands and fp32 accumulation whatever the dtype of its
        parameters (the reference's VAEConfig defaults to fp32 parameters, tld/configs.py:43): the kernels read packed bf16 /
        fp32 copies of the weights, exactly like the denoiser.  Only a module built with allow_aten=True keeps fp32 (or CPU)
        tensors on the ATen reference graph."""
        if w.is_cuda and not (self.allow_aten and w.dtype != torch.bfloat16):
            return torch.bfloat16
        return w.dtype

    def _packed(self, key: str, maker):
        """cache of repacked weights keyed on the parameter's storage / version / device"""
        w = self._p(key)
        cache = self.__dict__.setdefault("_wpack_cache", {})
        sig = (w.data_ptr(), w._version, w.device)
        ent = cache.get(key)
        if ent is None or ent[0] != sig:
            ent = (sig, maker(w.detach()))
            cache[key] = ent
        return ent[1]

    def _gemm_bias(self, t, wkey: str, bkey: str, n_pad: int = 0, k_pad: int = 0):
        """[M, K] bf16 rows (NHWC pixels / tokens) x nn.Linear-or-1x1-conv weight [N, K] + bias on the tcgen05 GEMM
        (tld_op_gemm, EPI_BIAS_BF16).  n_pad / k_pad: zero-pad the weight to that many outputs / inputs."""
        from . import _lib

        w = self._p(wkey)
        N, K = w.shape[0], w.shape[1]
        Np, Kp = max(N, n_pad), max(K, k_pad)

        def pack(wd):
            m = torch.zeros(Np, Kp, device=wd.device, dtype=torch.bfloat16)
            m[:N, :K] = wd.reshape(N, K).to(torch.bfloat16)
            return m

        wp = self._packed(wkey, pack)
        bp = self._packed(bkey, lambda b: torch.cat([b.float(), torch.zeros(Np - N, device=b.device)]).contiguous())
        M = t.shape[0]
        assert t.shape[1] == Kp and t.is_contiguous() and t.dtype == torch.bfloat16
        out = torch.empty(M, Np, device=t.device, dtype=torch.bfloat16)
        _lib.check(_lib.load().tld_op_gemm(1, t.data_ptr(), wp.data_ptr(), M, Np, Kp, out.data_ptr(), bp.data_ptr(),
                                           _lib.current_stream_ptr(t.device)), "tld_op_gemm")
        self.own_launches += 1
        return out

    @staticmethod
    def _rows(x):
        """NCHW-shaped channels_last tensor -> its memory as [B*H*W, C] rows"""
        B, Cc, H, W = x.shape
        return x.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1).reshape(B * H * W, Cc)

    @staticmethod
    def _from_rows(t, B, H, W):
        """[B*H*W, C] rows -> NCHW-shaped channels_last tensor (no copy)"""
        return t.view(B, H, W, t.shape[1]).permute(0, 3, 1, 2)

    def _on_kernels(self, x) -> bool:
        return x.is_cuda and x.dtype == torch.bfloat16 and not self.force_aten

    def _kernel_map_ok(self, x) -> bool:
        B, _, H, W = x.shape
        wb = min(W, 128)
        return (self._on_kernels(x) and (H * W) % 128 == 0 and 128 % wb == 0 and W % wb == 0 and H % (128 // wb) == 0)

    def _conv3x3_padded(self, x, name):
        """3x3 conv whose input and / or output channel count is below 64 (conv_in, conv_out): both are zero-padded to 64 and
        the tcgen05 implicit-GEMM kernel does the rest; the padded output channels are sliced off (they are exact zeros)."""
        from . import _lib

        w = self._p(name + ".weight")
        cout, cin = w.shape[0], w.shape[1]
        cinp, coutp = max(64, -(-cin // 64) * 64), max(64, -(-cout // 64) * 64)

        def pack(wd):
            m = torch.zeros(coutp, 3, 3, cinp, device=wd.device, dtype=torch.bfloat16)
            m[:cout, :, :, :cin] = wd.permute(0, 2, 3, 1).to(torch.bfloat16)
            return m.reshape(coutp, 9 * cinp).contiguous()

        wp = self._packed(name + ".weight", pack)
        bp = self._packed(name + ".bias", lambda b: torch.cat([b.float(), torch.zeros(coutp - cout, device=b.device)]).contiguous())
        B, _, H, W = x.shape
        if x.shape[1] != cinp:   # zero-pad the channels (NHWC memory)
            xp = torch.zeros((B, H, W, cinp), device=x.device, dtype=x.dtype).permute(0, 3, 1, 2)   # NHWC memory
            xp[:, :x.shape[1]] = x
            x = xp
        x = x.contiguous(memory_format=torch.channels_last)
        y = torch.empty((B, coutp, H, W), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        _lib.check(_lib.load().tld_vae_conv3x3(x.data_ptr(), wp.data_ptr(), bp.data_ptr(), y.data_ptr(), B, H, W, cinp, coutp,
                                               _lib.current_stream_ptr(x.device)), "tld_vae_conv3x3")
        self.own_launches += 1
        return y   # [B, coutp, H, W]; the caller keeps the padding (next layer is padded too) or slices [:, :cout]

    def _own_conv_ok(self, x, name) -> bool:
        """3x3 conv eligible for the tcgen05 implicit-GEMM kernel (tld_vae_conv3x3)"""
        w = self._p(name + ".weight")
        cout, cin, kh, kw = w.shape
        B, _, H, W = x.shape
        wb = min(W, 128)
        return (self._on_kernels(x) and kh == 3 and kw == 3 and cin % 64 == 0 and cout % 64 == 0
                and (H * W) % 128 == 0 and 128 % wb == 0 and W % wb == 0 and H % (128 // wb) == 0)

    def _conv(self, x, name, pad, bias=True, residual=None):
        """``residual`` (own 3x3 convolutions only): added in the GEMM epilogue before the rounding (ResnetBlock shortcut)."""
        if pad == 1 and self._own_conv_ok(x, name):
            from . import _lib

            w = self._p(name + ".weight")
            cout, cin = w.shape[0], w.shape[1]
            cache = self.__dict__.setdefault("_wpack_cache", {})
            sig = (w.data_ptr(), w._version, x.device)   # load_state_dict / .to() after a first decode must repack
            ent = cache.get(name)
            if ent is None or ent[0] != sig:
                # [Cout, Cin, 3, 3] -> [Cout, (ky, kx, cin)] bf16: the K order the implicit GEMM walks
                ent = (sig, w.detach().permute(0, 2, 3, 1).reshape(cout, 9 * cin).to(torch.bfloat16).contiguous())
                cache[name] = ent
            wp = ent[1]
            x = x.contiguous(memory_format=torch.channels_last)
            B, _, H, W = x.shape
            y = torch.empty((B, cout, H, W), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
            bp = self._f32(name + ".bias", x.device).data_ptr() if bias else None
            if bias and self.fuse_gn_stats and cout % GN_GROUPS == 0 and (cout // GN_GROUPS) % 4 == 0:
                # epilogue extras: the shortcut add, and the GroupNorm statistics partials of the stored output - whoever
                # normalises this tensor next (norm2, the next block's norm1, conv_norm_out) reads it once instead of twice
                part = torch.empty((B * H * W // 32, cout // 4, 2), device=x.device, dtype=torch.float32)
                rp = None
                if residual is not None:
                    residual = residual.contiguous(memory_format=torch.channels_last)
                    assert residual.shape == y.shape and residual.dtype == y.dtype
                    rp = residual.data_ptr()
                _lib.check(_lib.load().tld_vae_conv3x3_fused(x.data_ptr(), wp.data_ptr(), bp, y.data_ptr(), B, H, W, cin, cout, rp,
                                                             part.data_ptr(), _lib.current_stream_ptr(x.device)),
                           "tld_vae_conv3x3_fused")
                y._tld_gn_part = part   # consumed by _group_norm if it is handed exactly this tensor object
                self.own_launches += 1
                return y
            _lib.check(_lib.load().tld_vae_conv3x3(x.data_ptr(), wp.data_ptr(), bp, y.data_ptr(), B, H, W, cin, cout,
                                                   _lib.current_stream_ptr(x.device)), "tld_vae_conv3x3")
            self.own_launches += 1
            return y if residual is None else self._add_bias(residual, y)
        w = self._p(name + ".weight")
        if pad == 1 and bias and w.shape[2] == 3 and self._kernel_map_ok(x):          # thin 3x3: channels padded to 64
            return self._conv3x3_padded(x, name)[:, :w.shape[0]]
        if pad == 0 and bias and w.shape[2] == 1 and self._on_kernels(x) and w.shape[1] % 8 == 0 \
                and w.shape[0] % 32 == 0:                                              # 1x1 conv = GEMM over the pixels
            B, _, H, W = x.shape
            return self._from_rows(self._gemm_bias(self._rows(x), name + ".weight", name + ".bias"), B, H, W)
        self._aten(f"conv {name}", x)
        return F.conv2d(x, w, self._p(name + ".bias") if bias else None, padding=pad)

    def _f32(self, key: str, device) -> torch.Tensor:
        """fp32 copy of a 1-D parameter for the fused kernels (cached per device)"""
        cache = self.__dict__.setdefault("_f32_cache", {})
        p = self._p(key)
        sig = (p.data_ptr(), p._version, device)
        ent = cache.get(key)
        if ent is None or ent[0] != sig:
            ent = (sig, p.detach().to(device=device, dtype=torch.float32).contiguous())
            cache[key] = ent
        return ent[1]

    def _fusable(self, x) -> bool:
        Cc = x.shape[1]
        return (self._on_kernels(x) and Cc % 8 == 0 and Cc <= 512 and 256 % (Cc // 8) == 0 and (Cc // GN_GROUPS) % 4 == 0)

    def _group_norm(self, x, name, silu: bool, pre_bias=None):
        """act(GroupNorm(32, eps 1e-6)(x + pre_bias)); fused sm_100a kernels for bf16 channels-last CUDA tensors.
        ``pre_bias`` names the producing conv whose bias was left out of the conv call and is folded in here."""
        Cc = x.shape[1]
        if self._fusable(x):
            from . import _lib

            part = getattr(x, "_tld_gn_part", None)
            x = x.contiguous(memory_format=torch.channels_last)
            y = torch.empty_like(x)  # keeps the NHWC strides
            if part is not None and pre_bias is None and part.shape[0] * 32 == x.shape[0] * x.shape[2] * x.shape[3]:
                _lib.check(_lib.load().tld_vae_group_norm_from_conv(
                    x.data_ptr(), part.data_ptr(), self._f32(name + ".weight", x.device).data_ptr(),
                    self._f32(name + ".bias", x.device).data_ptr(), y.data_ptr(), x.shape[0], x.shape[2] * x.shape[3], Cc,
                    GN_GROUPS, GN_EPS, int(silu), _lib.current_stream_ptr(x.device)), "tld_vae_group_norm_from_conv")
                self.own_launches += 2  # finalize + apply
                return y
            pb = self._f32(pre_bias + ".bias", x.device).data_ptr() if pre_bias else None
            _lib.check(_lib.load().tld_vae_group_norm(
                x.data_ptr(), pb, self._f32(name + ".weight", x.device).data_ptr(),
                self._f32(name + ".bias", x.device).data_ptr(), y.data_ptr(), x.shape[0], x.shape[2] * x.shape[3], Cc,
                GN_GROUPS, GN_EPS, int(silu), _lib.current_stream_ptr(x.device)), "tld_vae_group_norm")
            self.own_launches += 2  # statistics + apply
            return y
        self._aten(f"group_norm {name}", x)
        if pre_bias:
            x = x + self._p(pre_bias + ".bias").view(1, -1, 1, 1)
        h = F.group_norm(x, GN_GROUPS, self._p(name + ".weight"), self._p(name + ".bias"), GN_EPS)
        return F.silu(h) if silu else h

    def _add_bias(self, x, h, bias_of=None):
        """x + h (+ the skipped bias of conv `bias_of`)"""
        if self._fusable(h) and x.shape == h.shape and x.dtype == h.dtype:
            from . import _lib

            x = x.contiguous(memory_format=torch.channels_last)
            h = h.contiguous(memory_format=torch.channels_last)
            out = torch.empty_like(h)
            bp = self._f32(bias_of + ".bias", h.device).data_ptr() if bias_of else None
            _lib.check(_lib.load().tld_vae_add_bias(x.data_ptr(), h.data_ptr(), bp, out.data_ptr(), h.numel(), h.shape[1],
                                                    _lib.current_stream_ptr(h.device)), "tld_vae_add_bias")
            self.own_launches += 1
            return out
        self._aten("residual add", h)
        return x + h + self._p(bias_of + ".bias").view(1, -1, 1, 1) if bias_of else x + h

    def _upsample2x(self, x):
        if self._on_kernels(x) and x.shape[1] % 8 == 0:
            from . import _lib

            x = x.contiguous(memory_format=torch.channels_last)
            B, Cc, H, W = x.shape
            y = torch.empty((B, Cc, 2 * H, 2 * W), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
            _lib.check(_lib.load().tld_vae_upsample2x(x.data_ptr(), y.data_ptr(), B, H, W, Cc,
                                                      _lib.current_stream_ptr(x.device)), "tld_vae_upsample2x")
            self.own_launches += 1
            return y
        self._aten("upsample", x)
        return F.interpolate(x, scale_factor=2.0, mode="nearest")

    def _resnet(self, x, name):
        if self._own_conv_ok(x, name + ".conv1"):
            # tcgen05 implicit-GEMM convs: the bias is part of the GEMM epilogue
            h = self._conv(self._group_norm(x, name + ".norm1", True), name + ".conv1", 1)
            if (name + ".conv_shortcut.weight") in self._layout:
                x = self._conv(x, name + ".conv_shortcut", 0)
            # the shortcut add rides in conv2's epilogue (and so do the statistics of the block's output for the next norm)
            return self._conv(self._group_norm(h, name + ".norm2", True), name + ".conv2", 1, residual=x)
        # library convs: their biases are folded into the next fused kernel (GroupNorm input / residual add)
        h = self._conv(self._group_norm(x, name + ".norm1", True), name + ".conv1", 1, bias=False)
        h = self._conv(self._group_norm(h, name + ".norm2", True, pre_bias=name + ".conv1"), name + ".conv2", 1, bias=False)
        if (name + ".conv_shortcut.weight") in self._layout:
            x = self._conv(x, name + ".conv_shortcut", 0)
        return self._add_bias(x, h, name + ".conv2")

    def _mid_attention(self, x, name):
        B, Cc, H, W = x.shape
        h = self._group_norm(x, name + ".group_norm", False)
        if self._on_kernels(x) and Cc % 32 == 0:
            # projections on the tcgen05 GEMM; the softmax(QK^T)V core of this single Cc-wide head: per image GEMM -> row
            # softmax -> GEMM (tld_vae_attention_core)
            from . import _lib

            t = self._rows(h)
            q, k, v = (self._gemm_bias(t, f"{name}.{p}.weight", f"{name}.{p}.bias") for p in ("to_q", "to_k", "to_v"))
            if (H * W) % 64 == 0 and Cc % 64 == 0:
                o = torch.empty_like(q)
                _lib.check(_lib.load().tld_vae_attention_core(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H * W, Cc,
                                                              _lib.current_stream_ptr(x.device)), "tld_vae_attention_core")
                self.own_launches += 3 * B
            else:
                self._aten(f"attention core {name}", x)
                o = F.scaled_dot_product_attention(*(u.view(B, 1, H * W, Cc) for u in (q, k, v))).reshape(B * H * W, Cc).contiguous()
            o = self._gemm_bias(o, name + ".to_out.0.weight", name + ".to_out.0.bias")
            return self._add_bias(x, self._from_rows(o, B, H, W))
        self._aten(f"attention {name}", x)
        t = h.permute(0, 2, 3, 1).reshape(B, H * W, Cc)
        q = F.linear(t, self._p(name + ".to_q.weight"), self._p(name + ".to_q.bias"))
        k = F.linear(t, self._p(name + ".to_k.weight"), self._p(name + ".to_k.bias"))
        v = F.linear(t, self._p(name + ".to_v.weight"), self._p(name + ".to_v.bias"))
        o = F.scaled_dot_product_attention(q.unsqueeze(1), k.unsqueeze(1), v.unsqueeze(1)).squeeze(1)  # 1 head
        o = F.linear(o, self._p(name + ".to_out.0.weight"), self._p(name + ".to_out.0.bias"))
        return x + o.reshape(B, H, W, Cc).permute(0, 3, 1, 2)

    def _decode_chunk(self, z):
        if self._kernel_map_ok(z):
            # stem on the library's kernels: latent channels zero-padded to 64 once; post_quant_conv (1x1) = one 64x64 GEMM
            # whose padded outputs are zeros again, which is exactly the padded input conv_in wants
            B, Cz, H, W = z.shape
            zp = torch.zeros((B, H, W, 64), device=z.device, dtype=z.dtype).permute(0, 3, 1, 2)   # NHWC memory
            zp[:, :Cz] = z
            t = self._gemm_bias(self._rows(zp), "post_quant_conv.weight", "post_quant_conv.bias", n_pad=64, k_pad=64)
            x = self._conv3x3_padded(self._from_rows(t, B, H, W), "decoder.conv_in")
        else:
            x = self._conv(z, "post_quant_conv", 0)
            x = self._conv(x, "decoder.conv_in", 1)
        x = self._resnet(x, "decoder.mid_block.resnets.0")
        x = self._mid_attention(x, "decoder.mid_block.attentions.0")
        x = self._resnet(x, "decoder.mid_block.resnets.1")
        n_up = len(self.block_out)
        for i in range(n_up):
            for j in range(LAYERS_PER_BLOCK + 1):
                x = self._resnet(x, f"decoder.up_blocks.{i}.resnets.{j}")
            if i != n_up - 1:
                x = self._upsample2x(x)
                x = self._conv(x, f"decoder.up_blocks.{i}.upsamplers.0.conv", 1)
        x = self._group_norm(x, "decoder.conv_norm_out", True)
        w = self._p("decoder.conv_out.weight")
        if self._on_kernels(x) and tuple(w.shape[:2]) == (3, 128):
            # 128 -> 3 channels at full resolution: HBM-bound direct convolution (tld_vae_conv_out3), fp32 NCHW image out
            from . import _lib

            wh = self._packed("decoder.conv_out.weight", lambda t: t.float().cpu().contiguous())   # host copies: kernel parameters
            bh = self._packed("decoder.conv_out.bias", lambda t: t.float().cpu().contiguous())
            x = x.contiguous(memory_format=torch.channels_last)
            B, _, H, W = x.shape
            y = torch.empty((B, 3, H, W), device=x.device, dtype=torch.float32)
            _lib.check(_lib.load().tld_vae_conv_out3(x.data_ptr(), wh.data_ptr(), bh.data_ptr(), y.data_ptr(), B, H, W,
                                                     _lib.current_stream_ptr(x.device)), "tld_vae_conv_out3")
            self.own_launches += 1
            return y
        return self._conv(x, "decoder.conv_out", 1)

    @torch.no_grad()
    def decode(self, z: torch.Tensor):
        """z [B,4,h,w] -> (image [B,3,8h,8w],) in z's dtype; processed in chunks of ``self.chunk`` images."""
        w = self._p("decoder.conv_in.weight")
        zz = z.to(device=w.device, dtype=self._compute_dtype(w))
        if zz.is_cuda:
            zz = zz.contiguous(memory_format=torch.channels_last)
        outs = [self._decode_chunk(zz[i:i + self.chunk]) for i in range(0, zz.shape[0], self.chunk)]
        return (torch.cat(outs).to(z.dtype),)

    @staticmethod
    def flops_per_image(latent_hw: int, block_out=BLOCK_OUT, latent_ch: int = LATENT_CH) -> float:
        """Algorithmic FLOPs of one decode (2*MACs of every conv/linear + the mid attention)."""
        top = block_out[-1]
        hw = latent_hw * latent_hw

        def res(n, cin, cout):
            return 2.0 * n * (cin * cout * 9 + cout * cout * 9 + (cin * cout if cin != cout else 0))

        f = 2.0 * hw * (latent_ch * latent_ch + latent_ch * top * 9)
        f += 2 * res(hw, top, top) + 4 * 2.0 * hw * top * top + 4.0 * hw * hw * top
        rev = list(reversed(block_out))
        prev, n = rev[0], hw
        for i, cout in enumerate(rev):
            for j in range(LAYERS_PER_BLOCK + 1):
                f += res(n, prev if j == 0 else cout, cout)
            if i != len(rev) - 1:
                n *= 4
                f += 2.0 * n * cout * cout * 9
            prev = cout
        f += 2.0 * n * block_out[0] * 3 * 9
        return f


# =====================================================================================================================
# Encoder half: ``vae.encode(x)[0].sample()`` as tld/data.py:35-41 calls it when the dataset's
# latents are produced.  Same status as the decoder: the architecture follows the published diffusers AutoencoderKL
# encoder with diffusers-compatible keys (``encoder.*``, ``quant_conv``) and PARITY IS UNPINNED; it is checked against
# ``oracle/vae_oracle.py:encode_moments`` on random weights.  All ResNet / GroupNorm / attention stages reuse the decoder's
# building blocks (tcgen05 implicit-GEMM 3x3 convs, fused GroupNorm+SiLU, fused residual add); the three stride-2
# downsampling convs, conv_in (3 channels), conv_out (8 channels) and quant_conv run on library kernels.
# =====================================================================================================================
def vae_encoder_param_layout(latent_ch: int = LATENT_CH, block_out=BLOCK_OUT, in_ch: int = 3) -> "Dict[str, tuple]":
    """Ordered {key: shape} of the encoder half of diffusers.AutoencoderKL (encoder + quant_conv)."""
    lay: Dict[str, tuple] = {}

    def conv(name, cout, cin, k):
        lay[name + ".weight"] = (cout, cin, k, k)
        lay[name + ".bias"] = (cout,)

    def norm(name, c):
        lay[name + ".weight"] = (c,)
        lay[name + ".bias"] = (c,)

    def resnet(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cout, cin, 3)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".conv_shortcut", cout, cin, 1)

    conv("encoder.conv_in", block_out[0], in_ch, 3)
    prev = block_out[0]
    for i, cout in enumerate(block_out):
        for j in range(LAYERS_PER_BLOCK):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", prev if j == 0 else cout, cout)
        if i != len(block_out) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
        prev = cout
    top = block_out[-1]
    resnet("encoder.mid_block.resnets.0", top, top)
    a = "encoder.mid_block.attentions.0"
    norm(a + ".group_norm", top)
    for p in ("to_q", "to_k", "to_v", "to_out.0"):
        lay[f"{a}.{p}.weight"] = (top, top)
        lay[f"{a}.{p}.bias"] = (top,)
    resnet("encoder.mid_block.resnets.1", top, top)
    norm("encoder.conv_norm_out", top)
    conv("encoder.conv_out", 2 * latent_ch, top, 3)
    conv("quant_conv", 2 * latent_ch, 2 * latent_ch, 1)
    return lay


class DiagonalGaussian:
    """The posterior ``encode`` returns (diffusers' DiagonalGaussianDistribution): ``mean``, ``logvar`` (clamped to
    [-30, 20]), ``std``; ``sample(generator)`` = mean + std * N(0, 1); ``mode()`` = mean."""

    def __init__(self, moments: torch.Tensor):
        self.mean, logvar = moments.chunk(2, dim=1)
        self.logvar = logvar.clamp(-30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None, noise: torch.Tensor = None) -> torch.Tensor:
        if noise is None:
            noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise.to(self.mean.dtype)

    def mode(self) -> torch.Tensor:
        return self.mean


class AutoencoderKLEncoder(AutoencoderKLDecoder):
    """``encode(x) -> (DiagonalGaussian,)`` with the SDXL-VAE encoder topology; x is an RGB image in [-1, 1]
    (tld/data.py:38-40).  Inherits the decoder's kernels-backed building blocks; only the parameter layout and the
    forward wiring differ."""

    def __init__(self, latent_ch: int = LATENT_CH, block_out: Tuple[int, ...] = BLOCK_OUT, chunk: int = 32,
                 allow_aten: bool = False):
        nn.Module.__init__(self)
        self.latent_ch, self.block_out, self.chunk = latent_ch, tuple(block_out), chunk
        self.allow_aten = allow_aten
        self.force_aten = False
        self.own_launches = 0
        self.fuse_gn_stats = True
        self._layout = vae_encoder_param_layout(latent_ch, block_out)
        for key, shape in self._layout.items():
            if key.endswith("weight") and len(shape) == 1:
                t = torch.ones(shape)
            elif key.endswith("bias"):
                t = torch.zeros(shape)
            else:
                fan_in = int(math.prod(shape[1:]))
                t = torch.empty(shape).uniform_(-1.0 / math.sqrt(fan_in), 1.0 / math.sqrt(fan_in))
            _attach(self, key, t)

    def _downsample(self, x, name):
        # diffusers Downsample2D(padding=0): pad right/bottom by one pixel, 3x3 conv with stride 2:
        #   out[y, x] = sum_k w[ky, kx] in[2y + ky, 2x + kx]  ==  the stride-1 'same' convolution sampled at (2y+1, 2x+1)
        # (its zero padding at the bottom / right edge is the same padded pixel) -> the tcgen05 kernel + a 2x sub-sampling
        if self._own_conv_ok(x, name):
            return self._conv(x, name, 1)[:, :, 1::2, 1::2].contiguous(memory_format=torch.channels_last)
        self._aten(f"downsample {name}", x)
        return F.conv2d(F.pad(x, (0, 1, 0, 1)), self._p(name + ".weight"), self._p(name + ".bias"), stride=2)

    def _encode_chunk(self, x):
        h = self._conv(x, "encoder.conv_in", 1)
        n_down = len(self.block_out)
        for i in range(n_down):
            for j in range(LAYERS_PER_BLOCK):
                h = self._resnet(h, f"encoder.down_blocks.{i}.resnets.{j}")
            if i != n_down - 1:
                h = self._downsample(h, f"encoder.down_blocks.{i}.downsamplers.0.conv")
        h = self._resnet(h, "encoder.mid_block.resnets.0")
        h = self._mid_attention(h, "encoder.mid_block.attentions.0")
        h = self._resnet(h, "encoder.mid_block.resnets.1")
        h = self._group_norm(h, "encoder.conv_norm_out", True)
        if self._kernel_map_ok(h) and self._own_conv_ok(h, "encoder.mid_block.resnets.1.conv2"):
            # conv_out (512 -> 8) with its outputs zero-padded to 64, quant_conv (1x1, 8 -> 8) as one 64x64 GEMM on those rows
            B, _, H, W = h.shape
            y = self._conv3x3_padded(h, "encoder.conv_out")
            t = self._gemm_bias(self._rows(y), "quant_conv.weight", "quant_conv.bias", n_pad=64, k_pad=64)
            return self._from_rows(t, B, H, W)[:, :2 * self.latent_ch]
        h = self._conv(h, "encoder.conv_out", 1)
        return self._conv(h, "quant_conv", 0)

    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = False):
        """x [B,3,H,W] in [-1,1] -> (DiagonalGaussian over [B,4,H/8,W/8],); processed in chunks of ``self.chunk`` images."""
        w = self._p("encoder.conv_in.weight")
        xx = x.to(device=w.device, dtype=self._compute_dtype(w))
        if xx.is_cuda:
            xx = xx.contiguous(memory_format=torch.channels_last)
        moments = torch.cat([self._encode_chunk(xx[i:i + self.chunk]) for i in range(0, xx.shape[0], self.chunk)])
        return (DiagonalGaussian(moments.to(x.dtype if x.is_floating_point() else w.dtype).contiguous()),)

    def decode(self, z):  # the encoder half has no decoder weights
        raise NotImplementedError("AutoencoderKLEncoder holds the encoder half only; use AutoencoderKLDecoder.decode")
