"""CLIP text tower — the text side of ``clip.load("ViT-L/14")[0]`` that the reference calls as
``clip_model.encode_text(clip.tokenize(prompt, truncate=True))`` (tld/diffusion.py:136-140,160,177).

``CLIPTextTower`` keeps the OpenAI checkpoint's text-side ``state_dict`` keys (``token_embedding.weight``,
``positional_embedding``, ``transformer.resblocks.{i}.{attn.in_proj_weight, attn.in_proj_bias, attn.out_proj.*, ln_1.*,
mlp.c_fc.*, mlp.c_proj.*, ln_2.*}``, ``ln_final.*``, ``text_projection``), so the text half of a real ViT-L/14 checkpoint loads
with ``load_state_dict(..., strict=True)`` once it is available (no network here: random init), and ``encode_text(tokens)`` runs
entirely on libtld_b200 kernels: the four linear layers of every block on the tcgen05 GEMM with its bias / bias + residual
epilogues, LayerNorm on ``tld_op_layernorm``, and the row-wise pieces of ``csrc/clip_kernels.cu`` (embedding gather, causal
attention over <= 128 tokens, QuickGELU, ln_final + EOT selection + text_projection).  fp32 residual stream, bf16 tensor-core
operands with fp32 accumulation - the same arithmetic contract as the denoiser.  No CPU path.

Not included: the BPE tokenizer (``clip.tokenize`` needs the vocabulary file that ships with the OpenAI package, which is not
installed here).  ``encode_text`` takes the int64 token ids ``clip.tokenize`` would produce ([B, 77], EOT = the highest id);
``DiffusionTransformer(..., text_encoder=lambda prompts: tower.encode_text(tokenize(prompts)))`` wires it in.

Parity: PINNED TO A THIRD-PARTY IMPLEMENTATION, not to the OpenAI package (absent): ``oracle/clip_oracle.py`` restates the
published ``CLIP.encode_text`` and is checked against ``transformers.CLIPTextModelWithProjection`` on random weights
(``tests/test_clip.py``); the kernels are checked against that oracle.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import _lib


class CLIPTextTower(nn.Module):
    def __init__(self, width: int = 768, layers: int = 12, vocab_size: int = 49408, context_length: int = 77, embed_dim: int = 768):
        super().__init__()
        if width % 64:
            raise ValueError("CLIPTextTower: width must be a multiple of 64 (heads = width / 64)")
        self.width, self.layers, self.vocab_size, self.context_length, self.embed_dim = width, layers, vocab_size, context_length, embed_dim
        self.own_launches = 0

        def P(name, t):
            mod = self
            *path, leaf = name.split(".")
            for part in path:
                if part not in mod._modules:
                    mod.add_module(part, nn.Module())
                mod = mod._modules[part]
            mod.register_parameter(leaf, nn.Parameter(t))

        # initialisation as in openai/CLIP model.py initialize_parameters
        P("token_embedding.weight", torch.randn(vocab_size, width) * 0.02)
        P("positional_embedding", torch.randn(context_length, width) * 0.01)
        proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
        attn_std, fc_std = width ** -0.5, (2 * width) ** -0.5
        for i in range(layers):
            b = f"transformer.resblocks.{i}."
            P(b + "attn.in_proj_weight", torch.randn(3 * width, width) * attn_std)
            P(b + "attn.in_proj_bias", torch.zeros(3 * width))
            P(b + "attn.out_proj.weight", torch.randn(width, width) * proj_std)
            P(b + "attn.out_proj.bias", torch.zeros(width))
            P(b + "ln_1.weight", torch.ones(width))
            P(b + "ln_1.bias", torch.zeros(width))
            P(b + "mlp.c_fc.weight", torch.randn(4 * width, width) * fc_std)
            P(b + "mlp.c_fc.bias", torch.zeros(4 * width))
            P(b + "mlp.c_proj.weight", torch.randn(width, 4 * width) * proj_std)
            P(b + "mlp.c_proj.bias", torch.zeros(width))
            P(b + "ln_2.weight", torch.ones(width))
            P(b + "ln_2.bias", torch.zeros(width))
        P("ln_final.weight", torch.ones(width))
        P("ln_final.bias", torch.zeros(width))
        P("text_projection", torch.randn(width, embed_dim) * width ** -0.5)

    # ------------------------------------------------------------------ packed copies (keyed on storage / version / device)
    def _p(self, key: str) -> torch.Tensor:
        mod = self
        *path, leaf = key.split(".")
        for name in path:
            mod = mod._modules[name]
        return mod._parameters[leaf]

    def _packed(self, key: str, dtype) -> torch.Tensor:
        w = self._p(key)
        cache = self.__dict__.setdefault("_pack_cache", {})
        sig = (w.data_ptr(), w._version, w.device, dtype)
        ent = cache.get(key)
        if ent is None or ent[0] != sig:
            ent = (sig, w.detach().to(dtype).contiguous())
            cache[key] = ent
        return ent[1]

    # ------------------------------------------------------------------ kernels
    def _ln(self, x, name):
        L = _lib.load()
        y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
        _lib.check(L.tld_op_layernorm(x.data_ptr(), self._packed(name + ".weight", torch.float32).data_ptr(),
                                      self._packed(name + ".bias", torch.float32).data_ptr(), y.data_ptr(), x.shape[0], x.shape[1],
                                      _lib.current_stream_ptr(x.device)), "tld_op_layernorm")
        self.own_launches += 1
        return y

    def _linear(self, a, wkey, bkey, residual=None):
        """a [M, K] bf16 @ W^T + b: bf16 output, or added into the fp32 `residual` (TMA reduce-add epilogue)"""
        L = _lib.load()
        w, bias = self._packed(wkey, torch.bfloat16), self._packed(bkey, torch.float32)
        M, K, N = a.shape[0], a.shape[1], w.shape[0]
        if residual is None:
            out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
            _lib.check(L.tld_op_gemm(1, a.data_ptr(), w.data_ptr(), M, N, K, out.data_ptr(), bias.data_ptr(),
                                     _lib.current_stream_ptr(a.device)), "tld_op_gemm")
        else:
            out = residual
            _lib.check(L.tld_op_gemm(2, a.data_ptr(), w.data_ptr(), M, N, K, out.data_ptr(), bias.data_ptr(),
                                     _lib.current_stream_ptr(a.device)), "tld_op_gemm")
        self.own_launches += 1
        return out

    @torch.no_grad()
    def encode_text(self, text: torch.Tensor) -> torch.Tensor:
        """text: int64 token ids [B, T <= context_length] (what clip.tokenize returns) -> fp32 [B, embed_dim] on the device"""
        dev = self._p("ln_final.weight").device
        if dev.type != "cuda":
            raise _lib.TldError("CLIPTextTower runs on CUDA (sm_100a) only; there is no CPU fallback")
        if text.dim() != 2 or text.shape[1] > self.context_length:
            raise ValueError(f"expected token ids [B, <= {self.context_length}], got {tuple(text.shape)}")
        L = _lib.load()
        ids = text.to(device=dev, dtype=torch.int64).contiguous()
        B, T = ids.shape
        W = self.width
        with torch.cuda.device(dev):
            st = _lib.current_stream_ptr(dev)
            x = torch.empty(B * T, W, device=dev, dtype=torch.float32)
            _lib.check(L.tld_clip_embed(ids.data_ptr(), self._packed("token_embedding.weight", torch.float32).data_ptr(),
                                        self._packed("positional_embedding", torch.float32).data_ptr(), x.data_ptr(), B, T, W,
                                        self.vocab_size, st), "tld_clip_embed")
            for i in range(self.layers):
                b = f"transformer.resblocks.{i}."
                qkv = self._linear(self._ln(x, b + "ln_1"), b + "attn.in_proj_weight", b + "attn.in_proj_bias")
                o = torch.empty(B * T, W, device=dev, dtype=torch.bfloat16)
                _lib.check(L.tld_clip_causal_attention(qkv.data_ptr(), o.data_ptr(), B, T, W, st), "tld_clip_causal_attention")
                self._linear(o, b + "attn.out_proj.weight", b + "attn.out_proj.bias", residual=x)
                h = self._linear(self._ln(x, b + "ln_2"), b + "mlp.c_fc.weight", b + "mlp.c_fc.bias")
                g = torch.empty_like(h)
                _lib.check(L.tld_clip_quick_gelu(h.data_ptr(), g.data_ptr(), h.numel(), st), "tld_clip_quick_gelu")
                self._linear(g, b + "mlp.c_proj.weight", b + "mlp.c_proj.bias", residual=x)
                self.own_launches += 2
            eot = ids.argmax(dim=-1).contiguous()      # the EOT token has the highest id (openai/CLIP encode_text)
            out = torch.empty(B, self.embed_dim, device=dev, dtype=torch.float32)
            _lib.check(L.tld_clip_final(x.data_ptr(), eot.data_ptr(), self._packed("ln_final.weight", torch.float32).data_ptr(),
                                        self._packed("ln_final.bias", torch.float32).data_ptr(),
                                        self._packed("text_projection", torch.float32).data_ptr(), out.data_ptr(), B, T, W,
                                        self.embed_dim, st), "tld_clip_final")
            self.own_launches += 2
        return out

    forward = encode_text
