"""Import the UNMODIFIED reference (`/root/reference/tld`) for golden generation.  TEST INFRA ONLY.

Works only in the build container (the GPU box has no `/root/reference`).  `tld.diffusion` needs
`clip` and `diffusers`, which are not installed; two name-only stub modules (no arithmetic) make
the import succeed so the reference's own sampling loop runs unmodified.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TLD_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "tld", "denoiser.py"))


def load_reference():
    """Return the reference modules as a namespace: .denoiser .transformer_blocks .diffusion .configs"""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "clip" not in sys.modules:
        sys.modules["clip"] = types.ModuleType("clip")
    if "diffusers" not in sys.modules:
        stub = types.ModuleType("diffusers")

        class AutoencoderKL:  # name only; the tests inject their own object with .decode()
            pass

        stub.AutoencoderKL = AutoencoderKL
        sys.modules["diffusers"] = stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import tld.configs  # noqa: F401
    import tld.denoiser  # noqa: F401
    import tld.diffusion  # noqa: F401
    import tld.transformer_blocks  # noqa: F401

    ns = types.SimpleNamespace(
        denoiser=sys.modules["tld.denoiser"],
        transformer_blocks=sys.modules["tld.transformer_blocks"],
        diffusion=sys.modules["tld.diffusion"],
        configs=sys.modules["tld.configs"],
    )
    assert os.path.realpath(ns.denoiser.__file__).startswith(os.path.realpath(REFERENCE_ROOT))
    return ns


class IdentityVAE:
    """Stands in for AutoencoderKL in the reference loop: decode(z) -> (z,) (diffusion.py:91)."""

    def decode(self, z):
        return (z,)
