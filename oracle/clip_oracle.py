"""CPU restatement of the text side of OpenAI CLIP (`model.encode_text`), the call the reference makes at
tld/diffusion.py:136-140 (`clip_model.encode_text(clip.tokenize(label, truncate=True))`).  TEST INFRASTRUCTURE ONLY.

The algorithm lives in a third-party dependency that is absent from /root/reference and from this image: `clip @
git+https://github.com/openai/CLIP.git` (pyproject.toml:38, unpinned HEAD), model ViT-L/14 (tld/configs.py: clip_model_name).
Published algorithm (openai/CLIP clip/model.py, `CLIP.encode_text` + `Transformer` / `ResidualAttentionBlock`):

    x = token_embedding(text) + positional_embedding                       # [B, 77, 768]
    for each of the 12 blocks:  x = x + attn(ln_1(x), causal mask);  x = x + c_proj(QuickGELU(c_fc(ln_2(x))))
    x = ln_final(x);  x = x[arange(B), text.argmax(-1)] @ text_projection  # features at the EOT token (highest id)

with nn.MultiheadAttention(768, 12) (in_proj_weight [2304, 768] = q | k | v, scale 1/8) and QuickGELU(x) = x sigmoid(1.702 x).
State-dict keys are the text-side keys of the OpenAI checkpoint.  PINNING: tests/test_clip.py checks this restatement against
`transformers.CLIPTextModelWithProjection` (transformers 5.5, an independent implementation of the same published model) on
random weights mapped key by key; the OpenAI package itself and its weights are unavailable offline.
"""
from __future__ import annotations

import torch


def text_layout(width: int = 768, layers: int = 12, vocab: int = 49408, n_ctx: int = 77, embed_dim: int = 768) -> "dict[str, tuple]":
    lay = {"token_embedding.weight": (vocab, width), "positional_embedding": (n_ctx, width)}
    for i in range(layers):
        b = f"transformer.resblocks.{i}."
        lay[b + "attn.in_proj_weight"] = (3 * width, width)
        lay[b + "attn.in_proj_bias"] = (3 * width,)
        lay[b + "attn.out_proj.weight"] = (width, width)
        lay[b + "attn.out_proj.bias"] = (width,)
        lay[b + "ln_1.weight"] = (width,)
        lay[b + "ln_1.bias"] = (width,)
        lay[b + "mlp.c_fc.weight"] = (4 * width, width)
        lay[b + "mlp.c_fc.bias"] = (4 * width,)
        lay[b + "mlp.c_proj.weight"] = (width, 4 * width)
        lay[b + "mlp.c_proj.bias"] = (width,)
        lay[b + "ln_2.weight"] = (width,)
        lay[b + "ln_2.bias"] = (width,)
    lay["ln_final.weight"] = (width,)
    lay["ln_final.bias"] = (width,)
    lay["text_projection"] = (width, embed_dim)
    return lay


def encode_text(sd: "dict[str, torch.Tensor]", text: torch.Tensor, heads: int = 12) -> torch.Tensor:
    """openai/CLIP clip/model.py CLIP.encode_text, fp32"""
    B, T = text.shape
    W = sd["positional_embedding"].shape[1]
    x = sd["token_embedding.weight"][text] + sd["positional_embedding"][:T]
    mask = torch.full((T, T), float("-inf")).triu_(1)            # build_attention_mask: causal
    i = 0
    while f"transformer.resblocks.{i}.ln_1.weight" in sd:
        b = f"transformer.resblocks.{i}."
        h = torch.nn.functional.layer_norm(x, (W,), sd[b + "ln_1.weight"], sd[b + "ln_1.bias"], 1e-5)
        qkv = h @ sd[b + "attn.in_proj_weight"].t() + sd[b + "attn.in_proj_bias"]
        q, k, v = (t.view(B, T, heads, W // heads).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        a = torch.softmax(q @ k.transpose(-1, -2) / (W // heads) ** 0.5 + mask, -1) @ v
        a = a.transpose(1, 2).reshape(B, T, W)
        x = x + a @ sd[b + "attn.out_proj.weight"].t() + sd[b + "attn.out_proj.bias"]
        h = torch.nn.functional.layer_norm(x, (W,), sd[b + "ln_2.weight"], sd[b + "ln_2.bias"], 1e-5)
        h = h @ sd[b + "mlp.c_fc.weight"].t() + sd[b + "mlp.c_fc.bias"]
        h = h * torch.sigmoid(1.702 * h)                         # QuickGELU
        x = x + h @ sd[b + "mlp.c_proj.weight"].t() + sd[b + "mlp.c_proj.bias"]
        i += 1
    x = torch.nn.functional.layer_norm(x, (W,), sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)
    return x[torch.arange(B), text.argmax(dim=-1)] @ sd["text_projection"]


def from_hf(hf_sd: "dict[str, torch.Tensor]") -> "dict[str, torch.Tensor]":
    """transformers.CLIPTextModelWithProjection state_dict -> OpenAI text-side keys"""
    sd = {"token_embedding.weight": hf_sd["text_model.embeddings.token_embedding.weight"],
          "positional_embedding": hf_sd["text_model.embeddings.position_embedding.weight"],
          "ln_final.weight": hf_sd["text_model.final_layer_norm.weight"], "ln_final.bias": hf_sd["text_model.final_layer_norm.bias"],
          "text_projection": hf_sd["text_projection.weight"].t().contiguous()}
    i = 0
    while f"text_model.encoder.layers.{i}.layer_norm1.weight" in hf_sd:
        s, b = f"text_model.encoder.layers.{i}.", f"transformer.resblocks.{i}."
        sd[b + "attn.in_proj_weight"] = torch.cat([hf_sd[s + f"self_attn.{p}_proj.weight"] for p in "qkv"])
        sd[b + "attn.in_proj_bias"] = torch.cat([hf_sd[s + f"self_attn.{p}_proj.bias"] for p in "qkv"])
        sd[b + "attn.out_proj.weight"], sd[b + "attn.out_proj.bias"] = hf_sd[s + "self_attn.out_proj.weight"], hf_sd[s + "self_attn.out_proj.bias"]
        sd[b + "ln_1.weight"], sd[b + "ln_1.bias"] = hf_sd[s + "layer_norm1.weight"], hf_sd[s + "layer_norm1.bias"]
        sd[b + "ln_2.weight"], sd[b + "ln_2.bias"] = hf_sd[s + "layer_norm2.weight"], hf_sd[s + "layer_norm2.bias"]
        sd[b + "mlp.c_fc.weight"], sd[b + "mlp.c_fc.bias"] = hf_sd[s + "mlp.fc1.weight"], hf_sd[s + "mlp.fc1.bias"]
        sd[b + "mlp.c_proj.weight"], sd[b + "mlp.c_proj.bias"] = hf_sd[s + "mlp.fc2.weight"], hf_sd[s + "mlp.fc2.bias"]
        i += 1
    return sd
