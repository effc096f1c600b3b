"""The reference's own test-suite (tests/test_diffuser.py in apapiu/transformer_latent_diffusion), re-run against this
package through `install_as_tld()`: same constructor calls, same arguments, same (shape / type / "runs") assertions.
Network-dependent pieces (HF VAE weights, CLIP) are replaced by random-init stand-ins exactly as SURVEY.md §4 notes."""
import os
from dataclasses import asdict

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tld():
    import transformer_latent_diffusion_b200 as pkg

    pkg.install_as_tld()
    import tld.configs
    import tld.denoiser
    import tld.diffusion

    return pkg


def test_denoiser_outputs(tld, num_imgs=4):  # reference tests/test_diffuser.py:27-46
    from tld.configs import DenoiserConfig, LTDConfig
    from tld.denoiser import Denoiser

    denoiser_cfg = DenoiserConfig(n_channels=4)
    cfg = LTDConfig(denoiser_cfg=denoiser_cfg)
    model = Denoiser(**asdict(denoiser_cfg)).cuda()
    img_size = denoiser_cfg.image_size
    x = torch.rand(num_imgs, denoiser_cfg.n_channels, img_size, img_size).cuda()
    noise_level = torch.rand(num_imgs, 1).cuda()
    label = torch.rand(num_imgs, cfg.denoiser_cfg.text_emb_size).cuda()
    assert sum(p.numel() for p in model.parameters()) == 868_800
    with torch.no_grad():
        output = model(x, noise_level, label)
    assert output.shape == torch.Size([num_imgs, denoiser_cfg.n_channels, img_size, img_size])


def test_diffusion_generator(tld):  # reference tests/test_diffuser.py:49-85 (random-init VAE instead of the HF download)
    from tld.configs import DenoiserConfig, LTDConfig
    from tld.denoiser import Denoiser
    from tld.diffusion import DiffusionGenerator
    from transformer_latent_diffusion_b200.vae import AutoencoderKLDecoder

    denoiser_cfg = DenoiserConfig(n_channels=4)
    cfg = LTDConfig(denoiser_cfg=denoiser_cfg)
    device = torch.device("cuda:0")
    denoiser = Denoiser(**asdict(denoiser_cfg)).to(cfg.denoiser_load.dtype)
    vae = AutoencoderKLDecoder().to(device=device, dtype=cfg.vae_cfg.vae_dtype)
    labels = torch.rand(1, cfg.denoiser_cfg.text_emb_size)
    diffuser = DiffusionGenerator(denoiser.to(device), vae, device, cfg.denoiser_load.dtype)
    out, latent = diffuser.generate(labels=labels, num_imgs=1, img_size=cfg.denoiser_cfg.image_size, class_guidance=3,
                                    seed=1, n_iter=5, exponent=1, scale_factor=8, sharp_f=0, bright_f=0)
    assert out.shape == (1, 3, 128, 128) and out.device.type == "cpu" and torch.isfinite(out).all()
    assert latent.shape == (1, 4, 16, 16) and latent.is_cuda


def test_full_generation_pipeline(tld):  # reference tests/test_diffuser.py:88-93 (CLIP replaced by a fake encoder)
    from PIL.Image import Image
    from tld.configs import DenoiserConfig, LTDConfig
    from tld.diffusion import DiffusionTransformer

    cfg = LTDConfig(denoiser_cfg=DenoiserConfig(n_channels=4))

    def fake_clip(prompts):
        g = torch.Generator().manual_seed(len(prompts[0]))
        return torch.randn(len(prompts), 768, generator=g)

    dt = DiffusionTransformer(cfg, text_encoder=fake_clip)
    out = dt.generate_image_from_text(prompt="a cute cat", num_imgs=4, n_iter=5)
    assert isinstance(out, Image) and out.size == (2 * 128 + 3 * 4, 2 * 128 + 3 * 4)


def test_training(tld, tmp_path):  # reference tests/test_diffuser.py:96-121
    from transformer_latent_diffusion_b200.configs import DataConfig, ModelConfig, TrainConfig
    from transformer_latent_diffusion_b200.train import main

    data_config = DataConfig(latent_path=str(tmp_path / "latents.npy"), text_emb_path=str(tmp_path / "text_emb.npy"),
                             val_path=str(tmp_path / "val_emb.npy"))
    model_cfg = ModelConfig(data_config=data_config,
                            train_config=TrainConfig(n_epoch=2, save_model=False, compile=False, use_wandb=False))
    n = 200
    img_size = model_cfg.denoiser_config.image_size
    np.save(data_config.latent_path, torch.randn(n, model_cfg.denoiser_config.n_channels, img_size, img_size).numpy())
    np.save(data_config.text_emb_path, torch.randn(n, model_cfg.denoiser_config.text_emb_size).numpy())
    np.save(data_config.val_path, torch.randn(8, model_cfg.denoiser_config.text_emb_size).numpy())
    ema = main(model_cfg, log_every=1)
    assert all(torch.isfinite(p).all() for p in ema.parameters())


def test_training_eval_checkpoint_resume(tld, tmp_path):
    """eval_gen + checkpoint save / resume of train.main (tld/train.py:23-40,92-102,140-158):
    the checkpoint has the reference's keys, eval grids are written on rank 0, and a from_scratch=False run picks up
    the EMA weights, the optimizer state and the step counter."""
    from PIL.Image import Image

    from transformer_latent_diffusion_b200.configs import DataConfig, ModelConfig, TrainConfig
    from transformer_latent_diffusion_b200.train import eval_gen, main
    from transformer_latent_diffusion_b200.vae import AutoencoderKLDecoder

    data_config = DataConfig(latent_path=str(tmp_path / "latents.npy"), text_emb_path=str(tmp_path / "text_emb.npy"),
                             val_path=str(tmp_path / "val_emb.npy"))
    ckpt = str(tmp_path / "ckpt.pt")
    tc = TrainConfig(n_epoch=1, batch_size=32, save_model=True, compile=False, use_wandb=False, model_name=ckpt,
                     save_and_eval_every_iters=3)
    model_cfg = ModelConfig(data_config=data_config, train_config=tc)
    dc = model_cfg.denoiser_config
    n = 160   # 5 steps of 32
    np.save(data_config.latent_path, torch.randn(n, dc.n_channels, dc.image_size, dc.image_size).numpy())
    np.save(data_config.text_emb_path, torch.randn(n, dc.text_emb_size).numpy())
    np.save(data_config.val_path, torch.randn(8, dc.text_emb_size).numpy())
    vae = AutoencoderKLDecoder().to(device="cuda", dtype=model_cfg.vae_cfg.vae_dtype)
    ema = main(model_cfg, log_every=1, vae=vae, eval_dir=str(tmp_path))
    assert main.last_global_step == 5
    assert os.path.exists(tmp_path / "eval_step0.png") and os.path.exists(tmp_path / "eval_step3.png")
    assert os.path.exists(tmp_path / "img.jpg")
    sd = torch.load(ckpt, map_location="cpu")
    assert set(sd) == {"model_ema", "opt_state", "global_step"} and sd["global_step"] == 3
    assert list(sd["model_ema"].keys()) == list(ema.state_dict().keys())
    # resume: the step counter continues and the run starts from the checkpoint's EMA weights
    tc2 = TrainConfig(n_epoch=1, batch_size=32, save_model=False, compile=False, use_wandb=False, model_name=ckpt,
                      from_scratch=False, lr=0.0, alpha=1.0)   # lr 0 and alpha 1: weights must stay the checkpoint's
    ema2 = main(ModelConfig(data_config=data_config, train_config=tc2), log_every=1)
    assert main.last_global_step == 3 + 5
    for k, v in ema2.state_dict().items():
        assert torch.equal(v.cpu(), sd["model_ema"][k]), k
    # eval_gen on its own: reference signature, PIL image of the 2 x 8 grid
    from transformer_latent_diffusion_b200.diffusion import DiffusionGenerator

    img = eval_gen(DiffusionGenerator(ema2.eval(), vae, torch.device("cuda"), torch.float32),
                   torch.randn(8, dc.text_emb_size, device="cuda"), dc.image_size, out_path=str(tmp_path / "g.png"))
    side = dc.image_size * 8
    assert isinstance(img, Image) and img.size == (8 * side + 9 * 4, 2 * side + 3 * 4)
