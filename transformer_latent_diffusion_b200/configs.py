"""Configuration dataclasses — API-identical to the reference's ``tld/configs.py`` (field names, order,
defaults), because callers build them by keyword and round-trip them through ``dataclasses.asdict``
(``Denoiser(**asdict(cfg.denoiser_cfg))``, tld/diffusion.py:145; ``asdict(config)`` tld/train.py:112).

Reference: /root/reference/tld/configs.py:4-92.  Only the shape of the data is shared with the reference;
the helpers below (`to_json` / `from_json`, validation of what the B200 kernels support) are new.
"""
from __future__ import annotations

import json
from dataclasses import asdict, dataclass, field, fields, is_dataclass
from typing import Optional

import torch

_DTYPES = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}


@dataclass
class DenoiserConfig:
    """Constructor arguments of ``Denoiser`` (reference tld/configs.py:21-31)."""

    image_size: int = 16          # latent height = width
    noise_embed_dims: int = 256   # width of the sinusoidal noise-level embedding
    patch_size: int = 2
    embed_dim: int = 128          # heads = embed_dim // 64
    dropout: float = 0
    n_layers: int = 3
    text_emb_size: int = 768      # CLIP ViT-L/14 pooled embedding
    n_channels: int = 4           # latent channels of the VAE
    mlp_multiplier: int = 4

    def check_b200_support(self) -> None:
        """Raise ValueError for shapes the sm_100a kernels do not cover (see DESIGN.md 'limits')."""
        if self.embed_dim % 64 or not 64 <= self.embed_dim <= 1024:
            raise ValueError("embed_dim must be a multiple of 64 in [64, 1024] (training: a multiple of 128)")
        if self.image_size % self.patch_size:
            raise ValueError("image_size must be divisible by patch_size")
        if ((self.image_size // self.patch_size) ** 2) % 64:
            raise ValueError("(image_size/patch_size)^2 must be a multiple of 64")
        if self.n_channels * self.patch_size ** 2 > 64:
            raise ValueError("n_channels*patch_size^2 must be <= 64")


@dataclass
class DenoiserLoad:
    """Checkpoint location + dtype for inference (reference tld/configs.py:33-37)."""

    dtype: torch.dtype = torch.float32
    file_url: Optional[str] = None
    local_filename: Optional[str] = None


@dataclass
class VaeConfig:
    """reference tld/configs.py:39-43"""

    vae_scale_factor: float = 8
    vae_name: str = "madebyollin/sdxl-vae-fp16-fix"
    vae_dtype: torch.dtype = torch.float32


@dataclass
class ClipConfig:
    """reference tld/configs.py:45-48"""

    clip_model_name: str = "ViT-L/14"
    clip_dtype: torch.dtype = torch.float16


@dataclass
class DataConfig:
    """Where the pre-encoded training data lives (reference tld/configs.py:50-55)."""

    latent_path: str
    text_emb_path: str
    val_path: str


@dataclass
class TrainConfig:
    """reference tld/configs.py:57-72 — beta_a/beta_b shape the training noise-level distribution."""

    batch_size: int = 128
    lr: float = 3e-4
    n_epoch: int = 100
    alpha: float = 0.999
    from_scratch: bool = True
    beta_a: float = 1
    beta_b: float = 2.5
    save_and_eval_every_iters: int = 1000
    run_id: str = ""
    model_name: str = ""
    compile: bool = True
    save_model: bool = True
    use_wandb: bool = True


@dataclass
class DataDownloadConfig:
    """Offline dataset preparation settings (reference tld/configs.py:4-19).  The data-prep pipeline itself
    (tld/data.py) is out of scope for this package; the dataclass is kept so ``ModelConfig`` round-trips."""

    data_link: str
    caption_col: str = "caption"
    url_col: str = "url"
    latent_save_path: str = "latents_folder"
    raw_imgs_save_path: str = "raw_imgs_folder"
    use_drive: bool = False
    initial_csv_path: str = "imgs.csv"
    number_sample_per_shard: int = 10000
    image_size: int = 256
    batch_size: int = 64
    download_data: bool = True
    first_n_rows: int = 1000000
    use_wandb: bool = False


@dataclass
class LTDConfig:
    """Inference configuration (reference tld/configs.py:75-81)."""

    denoiser_cfg: DenoiserConfig = field(default_factory=DenoiserConfig)
    denoiser_load: DenoiserLoad = field(default_factory=DenoiserLoad)
    vae_cfg: VaeConfig = field(default_factory=VaeConfig)
    clip_cfg: ClipConfig = field(default_factory=ClipConfig)


@dataclass
class ModelConfig:
    """Training configuration (reference tld/configs.py:84-92)."""

    data_config: DataConfig
    download_config: Optional[DataDownloadConfig] = None
    denoiser_config: DenoiserConfig = field(default_factory=DenoiserConfig)
    train_config: TrainConfig = field(default_factory=TrainConfig)
    vae_cfg: VaeConfig = field(default_factory=VaeConfig)
    clip_cfg: ClipConfig = field(default_factory=ClipConfig)


# ---- JSON round trip (README.md:120-122 of the reference suggests dumping configs to JSON) ------------------
def _encode(obj):
    if isinstance(obj, torch.dtype):
        return {"__dtype__": str(obj).replace("torch.", "")}
    raise TypeError(type(obj))


def to_json(cfg) -> str:
    return json.dumps(asdict(cfg), default=_encode, indent=1)


def _build(cls, data):
    kwargs = {}
    for f in fields(cls):
        if f.name not in data:
            continue
        v = data[f.name]
        if isinstance(v, dict) and "__dtype__" in v:
            v = _DTYPES[v["__dtype__"]]
        elif isinstance(v, dict):
            sub = {"denoiser_cfg": DenoiserConfig, "denoiser_config": DenoiserConfig, "denoiser_load": DenoiserLoad,
                   "vae_cfg": VaeConfig, "clip_cfg": ClipConfig, "data_config": DataConfig,
                   "download_config": DataDownloadConfig, "train_config": TrainConfig}[f.name]
            v = _build(sub, v)
        kwargs[f.name] = v
    return cls(**kwargs)


def from_json(cls, text: str):
    assert is_dataclass(cls)
    return _build(cls, json.loads(text))
