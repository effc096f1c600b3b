"""VAE decoder: the package module (library-kernel path) against the independent fp32 oracle restatement,
random weights (parity with diffusers is UNPINNED, see oracle/vae_oracle.py)."""
import math

import pytest
import torch

from conftest import rel_fro


def _small():
    from transformer_latent_diffusion_b200.vae import AutoencoderKLDecoder

    torch.manual_seed(0)
    m = AutoencoderKLDecoder(block_out=(32, 64, 64, 64), allow_aten=True)   # widths / CPU tensors the kernels do not take
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
@pytest.mark.parametrize("B,C,H", [(2, 128, 16), (3, 256, 8), (1, 512, 32), (2, 128, 64)])
def test_group_norm_silu_kernel(B, C, H):
    """tld_vae_group_norm (fused GroupNorm+SiLU, NHWC bf16) against torch fp32 on the same bf16 inputs."""
    from transformer_latent_diffusion_b200 import _lib

    g = torch.Generator(device="cuda").manual_seed(C + H)
    x = (torch.randn(B, C, H, H, device="cuda", generator=g) * 2 + 0.5).bfloat16().contiguous(memory_format=torch.channels_last)
    w = torch.randn(C, device="cuda", generator=g)
    b = torch.randn(C, device="cuda", generator=g)
    for silu in (0, 1):
        ref = torch.nn.functional.group_norm(x.float(), 32, w, b, 1e-6)
        ref = torch.nn.functional.silu(ref) if silu else ref
        y = torch.empty_like(x)
        _lib.check(_lib.load().tld_vae_group_norm(x.data_ptr(), None, w.data_ptr(), b.data_ptr(), y.data_ptr(), B, H * H, C, 32,
                                                  1e-6, silu, torch.cuda.current_stream().cuda_stream), "gn")
        assert y.is_contiguous(memory_format=torch.channels_last)
        assert rel_fro(y.float(), ref) < 4e-3
        pb = torch.randn(C, device="cuda", generator=g)
        ref2 = torch.nn.functional.group_norm(x.float() + pb.view(1, -1, 1, 1), 32, w, b, 1e-6)
        ref2 = torch.nn.functional.silu(ref2) if silu else ref2
        _lib.check(_lib.load().tld_vae_group_norm(x.data_ptr(), pb.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), B, H * H, C,
                                                  32, 1e-6, silu, torch.cuda.current_stream().cuda_stream), "gn")
        assert rel_fro(y.float(), ref2) < 4e-3
    h2 = torch.randn_like(x)
    out = torch.empty_like(x)
    _lib.check(_lib.load().tld_vae_add_bias(x.data_ptr(), h2.data_ptr(), b.data_ptr(), out.data_ptr(), x.numel(), C,
                                            torch.cuda.current_stream().cuda_stream), "add")
    assert rel_fro(out.float(), x.float() + h2.float() + b.view(1, -1, 1, 1)) < 4e-3
    up = torch.empty(B, C, 2 * H, 2 * H, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    _lib.check(_lib.load().tld_vae_upsample2x(x.data_ptr(), up.data_ptr(), B, H, H, C, torch.cuda.current_stream().cuda_stream), "up")
    assert torch.equal(up, torch.nn.functional.interpolate(x, scale_factor=2.0, mode="nearest"))


@pytest.mark.gpu
@pytest.mark.parametrize("ctas", [1, 2])
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 16, 128, 128), (1, 32, 32, 512, 256), (3, 64, 64, 64, 192),
                                            (1, 128, 128, 128, 64), (1, 8, 256, 64, 128)])
def test_conv3x3_implicit_gemm(B, H, W, Cin, Cout, ctas):
    """tld_vae_conv3x3 (tcgen05 implicit GEMM, 4-D TMA im2col) against torch conv2d in fp32 on the same bf16 data."""
    from transformer_latent_diffusion_b200 import _lib

    L = _lib.load()
    _lib.check(L.tld_set_option(b"gemm_ctas", ctas), "opt")
    try:
        g = torch.Generator(device="cuda").manual_seed(H + Cin)
        x = torch.randn(B, Cin, H, W, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
        w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (3 * Cin ** 0.5)).bfloat16()
        bias = torch.randn(Cout, device="cuda", generator=g)
        ref = torch.nn.functional.conv2d(x.float(), w.float(), bias, padding=1)
        wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
        y = torch.full((B, Cout, H, W), float("nan"), device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        _lib.check(L.tld_vae_conv3x3(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, W, Cin, Cout,
                                     torch.cuda.current_stream().cuda_stream), "conv")
        torch.cuda.synchronize()
        assert torch.isfinite(y.float()).all()
        assert rel_fro(y.float(), ref) < 4e-3
        # borders exercise the TMA zero fill
        assert rel_fro(y.float()[:, :, 0, :], ref[:, :, 0, :]) < 4e-3 and rel_fro(y.float()[:, :, :, -1], ref[:, :, :, -1]) < 4e-3
    finally:
        _lib.check(L.tld_set_option(b"gemm_ctas", 0), "opt")


@pytest.mark.gpu
def test_decode_gpu_fused_kernels_match_oracle():
    """block widths that take the fused GroupNorm/upsample kernels (channels/32 multiple of 4)"""
    from oracle import vae_oracle as V
    from transformer_latent_diffusion_b200.vae import AutoencoderKLDecoder

    torch.manual_seed(1)
    m = AutoencoderKLDecoder(block_out=(128, 128, 256, 256))      # allow_aten stays False: every layer must be a library kernel
    for k, p in m.named_parameters():
        if p.ndim == 1:
            p.data.add_(0.1 * torch.randn_like(p))
    z = torch.randn(2, 4, 16, 16)
    ref = V.decode({k: v.detach() for k, v in m.state_dict().items()}, z)
    (img,) = m.cuda().to(torch.bfloat16).decode(z.cuda())
    assert m.own_launches > 100
    assert rel_fro(img, ref) < 8e-2


@pytest.mark.gpu
def test_cuda_tensors_without_a_kernel_raise():
    """no silent ATen dispatch in the product path (north_star: no multi-backend dispatch): fp32 CUDA tensors, channel counts or
    map sizes the kernels do not cover raise unless the module was explicitly built with allow_aten=True"""
    from transformer_latent_diffusion_b200 import _lib
    from transformer_latent_diffusion_b200.vae import AutoencoderKLDecoder

    m = AutoencoderKLDecoder(block_out=(128, 128, 256, 256)).cuda()
    (a,) = m.decode(torch.randn(1, 4, 16, 16, device="cuda"))             # fp32 parameters: bf16 operands inside, fp32 out
    assert a.dtype == torch.float32 and m.own_launches > 50
    with pytest.raises(_lib.TldError):
        m.to(torch.bfloat16).decode(torch.randn(1, 4, 8, 8, device="cuda"))   # 64-pixel maps
    with pytest.raises(_lib.TldError):
        AutoencoderKLDecoder(block_out=(32, 64, 64, 64)).cuda().to(torch.bfloat16).decode(torch.randn(1, 4, 16, 16, device="cuda"))
    with pytest.raises(_lib.TldError):
        AutoencoderKLDecoder(block_out=(128, 128, 256, 256)).decode(torch.randn(1, 4, 16, 16))   # CPU


@pytest.mark.gpu
def test_decode_full_sdxl_width_vs_oracle_with_error_attribution():
    """The widths bench.py times - (128, 256, 512, 512), one 32x32 latent -> 256x256 image - against the fp32 oracle, with the
    error attributed: the SAME module run through torch's own bf16 kernels (force_aten) shows how much of the deviation is
    bf16 arithmetic through ~35 conv / GroupNorm layers with random weights and how much is ours.  PARITY IS UNPINNED (no
    diffusers here): both are measured against oracle/vae_oracle.py only."""
    from oracle import vae_oracle as V
    from transformer_latent_diffusion_b200.vae import AutoencoderKLDecoder

    torch.manual_seed(7)
    m = AutoencoderKLDecoder(allow_aten=True)
    for k, p in m.named_parameters():
        if p.ndim == 1:
            p.data.add_(0.1 * torch.randn_like(p))
    z = torch.randn(1, 4, 32, 32)
    ref = V.decode({k: v.detach() for k, v in m.state_dict().items()}, z)
    m = m.cuda().to(torch.bfloat16)
    m.allow_aten = False
    (mine,) = m.decode(z.cuda())
    m.allow_aten, m.force_aten = True, True
    (aten,) = m.decode(z.cuda())
    e_mine, e_aten = rel_fro(mine, ref), rel_fro(aten, ref)
    print(f"full-width decode rel-Fro vs fp32 oracle: libtld_b200 {e_mine:.3e}, torch bf16 kernels {e_aten:.3e}")
    assert e_mine < 3e-2          # measured 2.5e-2 (torch's own bf16 kernels: 3.4e-2): the round-1 bar of 3e-2 holds here
    assert e_mine < 1.5 * e_aten + 5e-3, (e_mine, e_aten)   # no worse than the stock bf16 graph: the deviation is bf16, not a kernel


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


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,nrow", [(1, 8, 8, 1), (4, 16, 24, 2), (5, 32, 32, 3), (16, 64, 64, 8)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_image_grid_uint8_matches_reference_postprocessing(B, H, W, nrow, dtype):
    """tld_image_grid_u8 == ToPILImage(make_grid((x + 1) / 2, nrow, padding=4).clip(0, 1)) bit for bit
    (tld/diffusion.py:185, tld/train.py:36) - the uint8 grid is built on the device."""
    from transformer_latent_diffusion_b200.diffusion import image_grid_uint8, make_image_grid

    g = torch.Generator().manual_seed(B * 100 + H)
    x = (torch.randn(B, 3, H, W, generator=g) * 0.8).to(dtype)
    ref = (make_image_grid(x, nrow=nrow, padding=4).permute(1, 2, 0) * 255).to(torch.uint8).numpy()
    got = image_grid_uint8(x.cuda(), nrow, 4)
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert (got == ref).all(), f"{(got != ref).sum()} of {got.size} bytes differ"


# ------------------------------------------------------------------------------------------------------------------
# VAE encode + uint8 latent (de)quantisers (tld/data.py:34-60)
# ------------------------------------------------------------------------------------------------------------------
import os

import numpy as np

GOLD_Q = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "latent_quant.npz")


def _small_encoder(block_out=(32, 64, 64, 64), seed=0):
    from transformer_latent_diffusion_b200.vae import AutoencoderKLEncoder

    torch.manual_seed(seed)
    m = AutoencoderKLEncoder(block_out=block_out, allow_aten=True)
    for k, p in m.named_parameters():
        if p.ndim == 1:
            p.data.add_(0.1 * torch.randn_like(p))
    return m


def test_encoder_layout_and_oracle_cpu():
    from oracle import vae_oracle as V
    from transformer_latent_diffusion_b200.vae import vae_encoder_param_layout

    lay = vae_encoder_param_layout()
    assert lay["encoder.conv_in.weight"] == (128, 3, 3, 3)
    assert lay["encoder.down_blocks.1.resnets.0.conv_shortcut.weight"] == (256, 128, 1, 1)
    assert "encoder.down_blocks.3.downsamplers.0.conv.weight" not in lay
    assert lay["encoder.conv_out.weight"] == (8, 512, 3, 3) and lay["quant_conv.weight"] == (8, 8, 1, 1)
    assert sum(math.prod(s) for s in lay.values()) == 34_163_592 + 72   # SDXL-VAE encoder + quant_conv
    m = _small_encoder()
    x = torch.rand(2, 3, 32, 32) * 2 - 1
    (post,) = m.encode(x)
    mean, logvar = V.encode_moments({k: v.detach() for k, v in m.state_dict().items()}, x)
    assert post.mean.shape == (2, 4, 4, 4)
    assert rel_fro(post.mean, mean) < 1e-4 and rel_fro(post.logvar, logvar) < 1e-4
    eps = torch.randn(2, 4, 4, 4)
    assert rel_fro(post.sample(noise=eps), V.gaussian_sample(mean, logvar, eps)) < 1e-4
    assert torch.equal(post.mode(), post.mean)


def test_oracle_latent_quantisers_match_reference_vectors():
    """oracle restatement of tld/data.py:51-60 == vectors produced by the reference's own source (make_golden_quant.py)"""
    from oracle import vae_oracle as V

    d = np.load(GOLD_Q)
    l32, l16 = torch.from_numpy(d["lat32"]), torch.from_numpy(d["lat16"])
    assert (V.quantize_latents(l32).numpy() == d["q32"]).all()
    assert (V.quantize_latents(l32, 5.0).numpy() == d["q32_clip5"]).all()
    assert (V.quantize_latents(l16).numpy() == d["q16"]).all()
    b = torch.arange(256, dtype=torch.uint8)
    assert (V.dequantize_latents(b).numpy() == d["deq"]).all() and (V.dequantize_latents(b, 5.0).numpy() == d["deq_clip5"]).all()


def test_data_helpers_refuse_cpu_tensors():
    from transformer_latent_diffusion_b200 import _lib
    from transformer_latent_diffusion_b200.data import dequantize_latents, quantize_latents

    with pytest.raises(_lib.TldError):
        quantize_latents(torch.zeros(4))
    with pytest.raises(_lib.TldError):
        dequantize_latents(torch.zeros(4, dtype=torch.uint8))


@pytest.mark.gpu
def test_latent_quantiser_kernels_bit_exact():
    """tld_latent_quantize / tld_latent_dequantize against the reference-generated vectors, every byte"""
    from transformer_latent_diffusion_b200.data import dequantize_latents, quantize_latents

    d = np.load(GOLD_Q)
    l32, l16 = torch.from_numpy(d["lat32"]).cuda(), torch.from_numpy(d["lat16"]).cuda()
    assert (quantize_latents(l32).cpu().numpy() == d["q32"]).all()
    assert (quantize_latents(l32, 5.0).cpu().numpy() == d["q32_clip5"]).all()
    assert (quantize_latents(l16).cpu().numpy() == d["q16"]).all()
    b = torch.arange(256, dtype=torch.uint8, device="cuda")
    got, got5 = dequantize_latents(b), dequantize_latents(b, 5.0)
    assert got.dtype == torch.float16
    assert (got.cpu().numpy().view(np.uint16) == d["deq"].view(np.uint16)).all()
    assert (got5.cpu().numpy().view(np.uint16) == d["deq_clip5"].view(np.uint16)).all()
    # size-independent properties at a dataset-sized tensor: idempotence of quantise(dequantise(.)) and monotonicity
    big = torch.randint(0, 256, (1 << 22,), dtype=torch.uint8, device="cuda")
    rt = quantize_latents(dequantize_latents(big)).int()
    assert (rt - big.int()).abs().max() <= 1   # fp16 rounding of the dequantised value may move a byte by one, never more
    x = torch.linspace(-25, 25, 100001, device="cuda")
    q = quantize_latents(x).int()
    assert (q[1:] >= q[:-1]).all() and q[0] == 0 and q[-1] == 255


@pytest.mark.gpu
def test_encode_gpu_matches_oracle_and_roundtrip():
    from oracle import vae_oracle as V
    from transformer_latent_diffusion_b200.data import decode_latents, dequantize_latents, encode_image, quantize_latents
    from transformer_latent_diffusion_b200.vae import AutoencoderKLDecoder

    m = _small_encoder(block_out=(128, 128, 256, 256), seed=3)   # widths that take the fused kernels / tcgen05 convs
    img = torch.rand(3, 3, 128, 128)
    mean, logvar = V.encode_moments({k: v.detach() for k, v in m.state_dict().items()}, img * 2 - 1)
    torch.backends.cudnn.allow_tf32 = False
    (p32,) = m.cuda().encode((img * 2 - 1).cuda())          # fp32 on the GPU = the ATen graph (allow_aten)
    assert rel_fro(p32.mean, mean) < 1e-4 and rel_fro(p32.logvar, logvar) < 1e-4
    mb = m.to(torch.bfloat16)
    mb.allow_aten = False                                   # bf16: every layer on the library's kernels or it raises
    (pb,) = mb.encode((img * 2 - 1).cuda())
    assert rel_fro(pb.mean, mean) < 8e-2
    assert mb.own_launches > 0, "the encoder did not reach the library kernels"
    # the data-preparation chain of tld/data.py: encode -> quantise (device) -> dequantise -> decode
    lat = encode_image(img, mb, generator=torch.Generator(device="cuda").manual_seed(0), to_cpu=False)
    assert lat.shape == (3, 4, 16, 16) and lat.dtype == torch.float16 and lat.is_cuda
    q = quantize_latents(lat)
    back = dequantize_latents(q)
    assert q.dtype == torch.uint8 and (back.float() - lat.float().clip(-20, 20)).abs().max() <= 40.0 / 255 + 1e-2
    dec = AutoencoderKLDecoder(block_out=(128, 128, 256, 256)).cuda().to(torch.bfloat16)
    out = decode_latents(back, dec)
    assert out.shape == (3, 3, 128, 128) and out.min() >= 0 and out.max() <= 1


def test_vae_caches_follow_load_state_dict():
    """weights loaded AFTER a first decode must be used (the packed / fp32 copies are keyed on the parameter's storage and
    version): a fresh module with the same weights gives the same image"""
    m = _small()
    z = torch.randn(1, 4, 8, 8)
    m.decode(z)
    donor = _small()
    with torch.no_grad():
        for p in donor.parameters():
            p.mul_(1.3)
    m.load_state_dict(donor.state_dict())
    (a,), (b,) = m.decode(z), donor.decode(z)
    assert torch.equal(a, b)


@pytest.mark.gpu
def test_vae_gpu_caches_follow_load_state_dict():
    """the same on the CUDA/bf16 path, where the repacked conv weights and fp32 norm parameters are cached"""
    from transformer_latent_diffusion_b200.vae import AutoencoderKLDecoder

    torch.manual_seed(5)
    m = AutoencoderKLDecoder(block_out=(128, 128, 256, 256)).cuda().to(torch.bfloat16)
    donor = AutoencoderKLDecoder(block_out=(128, 128, 256, 256)).cuda().to(torch.bfloat16)
    z = torch.randn(2, 4, 16, 16, device="cuda")
    (before,) = m.decode(z)          # fills the caches with m's initial weights
    m.load_state_dict(donor.state_dict())
    (a,), (b,) = m.decode(z), donor.decode(z)
    # same weights -> same image up to library-kernel run-to-run differences; stale caches would give `before` again
    # (two instances go through cuDNN/cuBLAS for conv_in/conv_out/attention: algorithm choices differ at the bf16 level,
    #  a few 1e-3 .. 1e-2 after 30 layers; unrelated weights differ by O(1))
    assert rel_fro(a, b) < 5e-2 and rel_fro(a, before) > 0.3


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W", [(1, 8, 32), (2, 64, 64), (3, 40, 72), (1, 256, 256)])
def test_conv_out3_direct_kernel(B, H, W):
    """decoder.conv_out (128 -> 3, 3x3 'same') as the HBM-bound direct convolution with constant-bank weights, incl. ragged tiles"""
    from transformer_latent_diffusion_b200 import _lib

    g = torch.Generator(device="cuda").manual_seed(B * H + W)
    x = torch.randn(B, 128, H, W, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    w = torch.randn(3, 128, 3, 3, device="cuda", generator=g) / 34.0
    bias = torch.randn(3, device="cuda", generator=g)
    # float64 reference: an fp32 cuDNN convolution may run in TF32 (1e-4) depending on the algorithm the heuristic picks
    ref = torch.nn.functional.conv2d(x.double(), w.double(), bias.double(), padding=1).float()
    out = torch.full((B, 3, H, W), float("nan"), device="cuda")
    wh, bh = w.cpu().contiguous(), bias.cpu().contiguous()
    _lib.check(_lib.load().tld_vae_conv_out3(x.data_ptr(), wh.data_ptr(), bh.data_ptr(), out.data_ptr(), B, H, W,
                                             torch.cuda.current_stream().cuda_stream), "conv_out3")
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert rel_fro(out, ref) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,cin,cout", [(2, 16, 16, 128, 128), (1, 32, 32, 256, 128), (3, 16, 32, 128, 256), (1, 16, 16, 512, 512)])
def test_conv3x3_fused_residual_and_groupnorm_partials(B, H, W, cin, cout):
    """conv + bias + shortcut add in the GEMM epilogue, and the GroupNorm partials of the stored output (the consumer reads
    the tensor once): against fp32 torch, and the normalisation against the two-pass kernel on the same tensor"""
    from transformer_latent_diffusion_b200 import _lib

    L = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(B * 7 + cin + cout)
    x = torch.randn(B, cin, H, W, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    res = torch.randn(B, cout, H, W, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cin ** 0.5)).bfloat16()
    bias = torch.randn(cout, device="cuda", generator=g)
    wp = w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    y = torch.empty_like(res)
    part = torch.full((B * H * W // 32, cout // 4, 2), float("nan"), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.tld_vae_conv3x3_fused(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, W, cin, cout,
                                       res.data_ptr(), part.data_ptr(), st), "conv3x3_fused")
    ref = torch.nn.functional.conv2d(x.double(), w.double(), bias.double(), padding=1) + res.double()
    assert rel_fro(y.double(), ref) < 4e-3
    # partials: per (image, group) sums of the STORED values
    assert torch.isfinite(part).all()
    cpg = cout // 32
    ps = part.view(B, H * W // 32, 32, cpg // 4, 2).double().sum(dim=(1, 3))            # [B, 32 groups, 2]
    yv = y.double().view(B, 32, cpg, H * W)
    assert torch.allclose(ps[..., 0], yv.sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(ps[..., 1], (yv * yv).sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)
    # the normalisation from the partials == the two-pass kernel on the same tensor
    gamma = torch.randn(cout, device="cuda", generator=g) * 0.2 + 1
    beta = torch.randn(cout, device="cuda", generator=g) * 0.2
    a, b2 = torch.empty_like(y), torch.empty_like(y)
    _lib.check(L.tld_vae_group_norm_from_conv(y.data_ptr(), part.data_ptr(), gamma.data_ptr(), beta.data_ptr(), a.data_ptr(), B,
                                              H * W, cout, 32, 1e-6, 1, st), "gn_from_conv")
    _lib.check(L.tld_vae_group_norm(y.data_ptr(), None, gamma.data_ptr(), beta.data_ptr(), b2.data_ptr(), B, H * W, cout, 32, 1e-6,
                                    1, st), "gn")
    torch.cuda.synchronize()
    assert rel_fro(a.float(), b2.float()) < 2e-3
    # without residual / partials the plain entry point gives the same convolution
    y2 = torch.empty_like(y)
    _lib.check(L.tld_vae_conv3x3_fused(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), y2.data_ptr(), B, H, W, cin, cout, None, None,
                                       st), "conv3x3_fused")
    y3 = torch.empty_like(y)
    _lib.check(L.tld_vae_conv3x3(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), y3.data_ptr(), B, H, W, cin, cout, st), "conv3x3")
    torch.cuda.synchronize()
    assert torch.equal(y2, y3)


@pytest.mark.gpu
@pytest.mark.parametrize("B,n,C", [(2, 256, 128), (3, 1024, 512), (1, 64, 64)])
def test_vae_attention_core(B, n, C):
    """mid-block attention core (one C-wide head) on the library's GEMM + row softmax vs fp32 torch"""
    from transformer_latent_diffusion_b200 import _lib

    g = torch.Generator(device="cuda").manual_seed(n + C)
    q, k, v = (torch.randn(B, n, C, device="cuda", generator=g).bfloat16() for _ in range(3))
    q = q * 2.0
    o = torch.empty_like(q)
    _lib.check(_lib.load().tld_vae_attention_core(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, n, C,
                                                  torch.cuda.current_stream().cuda_stream), "vae_attention")
    ref = torch.softmax(q.double() @ k.double().transpose(1, 2) / C ** 0.5, -1) @ v.double()
    assert rel_fro(o.double(), ref) < 6e-3
