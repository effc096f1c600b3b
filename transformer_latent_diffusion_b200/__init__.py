"""transformer_latent_diffusion_b200 — B200-native (sm_100a) hot path of apapiu/transformer_latent_diffusion.

Drop-in for the reference's Python API on that path:

    tld.denoiser.Denoiser                  -> transformer_latent_diffusion_b200.denoiser.Denoiser
    tld.diffusion.DiffusionGenerator       -> transformer_latent_diffusion_b200.diffusion.DiffusionGenerator
    tld.diffusion.DiffusionTransformer     -> transformer_latent_diffusion_b200.diffusion.DiffusionTransformer
    tld.configs.*                          -> transformer_latent_diffusion_b200.configs.*
    tld.train.main                         -> transformer_latent_diffusion_b200.train.main

``install_as_tld()`` registers the package under the name ``tld`` so unmodified caller code
(``from tld.denoiser import Denoiser``) picks up this implementation.
"""
from __future__ import annotations

import sys

__version__ = "0.1.0"


def install_as_tld() -> None:
    """Alias this package as ``tld`` in ``sys.modules`` (refuses if a different ``tld`` is already imported)."""
    import importlib

    me = sys.modules[__name__]
    cur = sys.modules.get("tld")
    if cur is not None and cur is not me:
        raise RuntimeError("another module named 'tld' is already imported")
    sys.modules["tld"] = me
    for sub in ("configs", "denoiser", "diffusion", "train"):
        sys.modules[f"tld.{sub}"] = importlib.import_module(f"{__name__}.{sub}")
