#!/usr/bin/env python3
"""Golden vectors for the uint8 latent (de)quantisers (tld/data.py:51-60), produced by the reference's OWN source.

`tld.data` cannot be imported offline (h5py / webdataset / img2dataset / clip are missing), so the two function
definitions are cut out of /root/reference/tld/data.py with `ast` and executed unmodified in a namespace that only
provides `torch` and `Tensor`.  Run in the build container:  python tests/golden/make_golden_quant.py
"""
import ast
import os

import numpy as np
import torch

SRC = "/root/reference/tld/data.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "latent_quant.npz")


def reference_functions():
    tree = ast.parse(open(SRC).read())
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("quantize_latents", "dequantize_latents")]
    assert len(wanted) == 2
    ns = {"torch": torch, "Tensor": torch.Tensor}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), SRC, "exec"), ns)
    return ns["quantize_latents"], ns["dequantize_latents"]


def main():
    quant, dequant = reference_functions()
    g = torch.Generator().manual_seed(77)
    lat32 = torch.randn(4, 4, 16, 16, generator=g) * 9.0
    lat32.view(-1)[:8] = torch.tensor([-25.0, -20.0, -19.999, 0.0, 19.999, 20.0, 25.0, 1e-3])   # clip edges
    lat16 = lat32.to(torch.float16)                       # what encode_image hands over (fp16, tld/data.py:36-41)
    all_bytes = torch.arange(256, dtype=torch.uint8)
    np.savez(OUT,
             lat32=lat32.numpy(), q32=quant(lat32).numpy(), q32_clip5=quant(lat32, clip_val=5.0).numpy(),
             lat16=lat16.numpy(), q16=quant(lat16).numpy(),
             deq=dequant(all_bytes).numpy(), deq_clip5=dequant(all_bytes, clip_val=5.0).numpy())
    print("wrote", OUT)


if __name__ == "__main__":
    main()
