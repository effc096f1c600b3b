"""CLIP text tower (reference call site tld/diffusion.py:136-140).  The OpenAI `clip` package and its
weights are absent: the oracle restatement (oracle/clip_oracle.py) is pinned against transformers.CLIPTextModelWithProjection (an
independent implementation of the same published model) and the kernels are checked against the oracle."""
import pytest
import torch

from conftest import rel_fro


def _tokens(B, T, vocab, g):
    ids = torch.randint(1, vocab - 10, (B, T), generator=g)
    for b in range(B):
        e = int(torch.randint(2, T, (1,), generator=g))
        ids[b, e] = vocab - 1          # EOT = the highest id
        ids[b, e + 1:] = 0
    return ids


def test_oracle_matches_transformers_clip_text_model():
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection

    from oracle import clip_oracle as CO

    torch.manual_seed(0)
    cfg = CLIPTextConfig(vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=3, num_attention_heads=2,
                         max_position_embeddings=77, hidden_act="quick_gelu", projection_dim=96, eos_token_id=999,
                         bos_token_id=998, pad_token_id=0)
    m = CLIPTextModelWithProjection(cfg).eval()
    for p in m.parameters():
        p.data.add_(0.05 * torch.randn_like(p))
    ids = _tokens(4, 77, 1000, torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = m(input_ids=ids).text_embeds
    out = CO.encode_text(CO.from_hf(m.state_dict()), ids, heads=2)
    assert rel_fro(out, ref) < 1e-5


def test_tower_state_dict_keys_are_the_openai_text_keys():
    from oracle import clip_oracle as CO
    from transformer_latent_diffusion_b200.clip_text import CLIPTextTower

    t = CLIPTextTower(width=128, layers=2, vocab_size=500, embed_dim=64)
    lay = CO.text_layout(width=128, layers=2, vocab=500, embed_dim=64)
    sd = t.state_dict()
    assert set(sd.keys()) == set(lay.keys())
    assert all(tuple(sd[k].shape) == lay[k] for k in lay)
    full = CO.text_layout()     # ViT-L/14 text side
    assert sum(int(torch.tensor(s).prod()) for s in full.values()) == 123_650_304 - 0   # = transformers' CLIPTextModelWithProjection


def test_tower_refuses_cpu():
    from transformer_latent_diffusion_b200 import _lib
    from transformer_latent_diffusion_b200.clip_text import CLIPTextTower

    with pytest.raises(_lib.TldError):
        CLIPTextTower(width=64, layers=1, vocab_size=100, embed_dim=32).encode_text(torch.zeros(1, 77, dtype=torch.long))


@pytest.mark.gpu
@pytest.mark.parametrize("width,layers,B,T,P", [(128, 2, 3, 77, 96), (768, 12, 2, 77, 768), (512, 3, 5, 40, 512)])
def test_encode_text_matches_oracle(width, layers, B, T, P):
    from oracle import clip_oracle as CO
    from transformer_latent_diffusion_b200.clip_text import CLIPTextTower

    torch.manual_seed(width + layers)
    t = CLIPTextTower(width=width, layers=layers, vocab_size=2000, context_length=77, embed_dim=P)
    with torch.no_grad():
        for k, p in t.named_parameters():
            if p.ndim == 1:
                p.add_(0.1 * torch.randn_like(p))
    ids = _tokens(B, T, 2000, torch.Generator().manual_seed(7))
    sd = {k: v.detach().clone() for k, v in t.state_dict().items()}
    with torch.no_grad():
        ref = CO.encode_text(sd, ids, heads=width // 64)
    out = t.cuda().encode_text(ids.cuda())
    assert out.shape == (B, P) and out.dtype == torch.float32 and out.is_cuda
    assert t.own_launches > 0
    err = rel_fro(out, ref)
    assert err < 1e-2, f"rel_fro={err:.3e}"      # bf16 operands / fp32 accumulate through `layers` blocks, as the denoiser


@pytest.mark.gpu
def test_text_to_image_wiring_with_the_tower():
    """DiffusionTransformer.generate_image_from_text end to end on library kernels: CLIP tower -> sampler -> VAE -> uint8 grid"""
    from transformer_latent_diffusion_b200.clip_text import CLIPTextTower
    from transformer_latent_diffusion_b200.configs import DenoiserConfig, DenoiserLoad, LTDConfig
    from transformer_latent_diffusion_b200.diffusion import DiffusionTransformer

    torch.manual_seed(0)
    tower = CLIPTextTower(layers=2).cuda()
    g = torch.Generator().manual_seed(3)

    def encoder(prompts):
        return tower.encode_text(_tokens(len(prompts), 77, 49408, g))     # stand-in for clip.tokenize (vocabulary file absent)

    cfg = LTDConfig(denoiser_cfg=DenoiserConfig(image_size=16, n_layers=1), denoiser_load=DenoiserLoad(file_url=None, local_filename=None))
    dt = DiffusionTransformer(cfg, text_encoder=encoder, device=torch.device("cuda:0"))
    img = dt.generate_image_from_text("a cute grey cat", num_imgs=1, n_iter=4)
    assert img.size == (128, 128)
