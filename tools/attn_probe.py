"""Time tld_op_self_attention alone (CUDA events, L2 flushed between reps).  python tools/attn_probe.py --tok 4096 --batch 8"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_latent_diffusion_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--tok", type=int, default=256)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--impl", type=int, default=0)
ap.add_argument("--exp-emu", type=int, default=-1)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
L = _lib.load()
if a.exp_emu >= 0:
    _lib.check(L.tld_set_option(b"attention_exp_emu", a.exp_emu), "opt")
T = a.batch * a.tok
qkv = torch.randn(T, 3 * a.dim, device="cuda").bfloat16()
x = torch.zeros(T, a.dim, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
ts = []
for i in range(a.reps + 2):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(L.tld_op_self_attention(_lib.ptr(qkv), _lib.ptr(x), a.batch, a.tok, a.dim, a.impl, st), "attn")
    e1.record()
    torch.cuda.synchronize()
    if i >= 2:
        ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
flops = 4.0 * a.batch * (a.dim // 64) * a.tok * a.tok * 64
print(f"attention impl={a.impl} emu={a.exp_emu} B={a.batch} tok={a.tok} D={a.dim}: median {ts[len(ts)//2]:.1f} us  min {ts[0]:.1f} us"
      f"  -> {flops / ts[len(ts)//2] / 1e6:.1f} TFLOP/s")
