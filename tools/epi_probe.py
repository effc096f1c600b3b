"""Epilogue-only cost of the GEMM kernel: M=32768, N=768 with a tiny K (64), per epilogue kind.  Shows what the
TMA reduce-add into the fp32 residual stream costs against a plain TMA store of the same bytes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_latent_diffusion_b200 import _lib
L = _lib.load()
M, N = 32768, 768
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for K in (64, 768):
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.zeros(N, device="cuda")
    for epi, name, dt in ((0, "bf16 store", torch.bfloat16), (4, "f32 store", torch.float32), (2, "f32 reduce-add", torch.float32)):
        out = torch.zeros(M, N, device="cuda", dtype=dt)
        ts = []
        for i in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(L.tld_op_gemm(epi, _lib.ptr(A), _lib.ptr(W), M, N, K, _lib.ptr(out), _lib.ptr(bias) if epi == 2 else None, st), "g")
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        print(f"K={K:4d} {name:15s}: {ts[len(ts)//2]:7.1f} us   out bytes {out.numel()*out.element_size()/1e6:.0f} MB")
