"""VAE decoder: the package module (library-kernel path) against the independent fp32 oracle restatement,
random weights (parity with diffusers is UNPINNED, see oracle/vae_oracle.py)."""
import math

import pytest
import torch

from conftest import rel_fro


def _small():
    from transformer_latent_diffusion_b200.vae import AutoencoderKLDecoder

    torch.manual_seed(0)
    m = AutoencoderKLDecoder(block_out=(32, 64, 64, 64))
    for k, p in m.named_parameters():  # exercise affine/bias terms
        if p.ndim == 1:
            p.data.add_(0.1 * torch.randn_like(p))
    return m


def test_layout_keys_and_flops():
    from transformer_latent_diffusion_b200.vae import AutoencoderKLDecoder, vae_param_layout

    lay = vae_param_layout()
    assert lay["decoder.conv_in.weight"] == (512, 4, 3, 3)
    assert lay["decoder.up_blocks.2.resnets.0.conv_shortcut.weight"] == (256, 512, 1, 1)
    assert lay["decoder.up_blocks.3.resnets.0.conv_shortcut.weight"] == (128, 256, 1, 1)
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in lay
    assert lay["decoder.mid_block.attentions.0.to_out.0.weight"] == (512, 512)
    assert lay["decoder.conv_out.weight"] == (3, 128, 3, 3)
    n = sum(math.prod(s) for s in lay.values())
    assert n == 49_490_199, n  # SDXL-VAE decoder (49,490,179) + post_quant_conv (20)
    # ~0.62 TFLOP per 256-px image
    assert abs(AutoencoderKLDecoder.flops_per_image(32) / 1e12 - 0.62) < 0.03


def test_decode_matches_oracle_cpu():
    from oracle import vae_oracle as V

    m = _small()
    z = torch.randn(2, 4, 8, 8)
    (img,) = m.decode(z)
    ref = V.decode({k: v.detach() for k, v in m.state_dict().items()}, z)
    assert img.shape == (2, 3, 64, 64)
    assert rel_fro(img, ref) < 1e-4


@pytest.mark.gpu
def test_decode_gpu_bf16_matches_oracle():
    from oracle import vae_oracle as V

    m = _small()
    z = torch.randn(3, 4, 16, 16)
    ref = V.decode({k: v.detach() for k, v in m.state_dict().items()}, z)
    torch.backends.cudnn.allow_tf32 = False
    (img32,) = m.cuda().decode(z.cuda())
    assert rel_fro(img32, ref) < 1e-4
    # bf16 weights AND activations through ~35 conv/GroupNorm layers with random weights: a few percent
    (img,) = m.to(torch.bfloat16).decode(z.cuda())
    assert img.dtype == torch.float32 and rel_fro(img, ref) < 8e-2
