"""Developer tool: timeline of CTA 0 of attention_tc2 (library must be built with -DTLD_TRACE on attention_tc2.cu)."""
import argparse, ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_latent_diffusion_b200 import _lib
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--tok", type=int, default=256)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--events", type=int, default=120)
a = ap.parse_args()
L = _lib.load()
T = a.batch * a.tok
qkv = torch.randn(T, 3 * a.dim, device="cuda").bfloat16()
x = torch.zeros(T, a.dim, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(L.tld_op_self_attention(_lib.ptr(qkv), _lib.ptr(x), a.batch, a.tok, a.dim, 3, st), "attn")
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (3 * 2048))()
raw = ctypes.CDLL(L._name)
raw.tld_debug_attention_trace.argtypes = [ctypes.c_void_p]
assert raw.tld_debug_attention_trace(buf) == 0
names = {0: {0: "load Q", 1: "load K", 2: "load V"}, 1: {0: "q_full", 1: "issue S", 2: "issue PV"},
         2: {0: "tile start", 1: "s_full", 2: "p_arrive", 3: "o_full(last)", 4: "tile end"}}
ev = []
for role in range(3):
    for i in range(2048):
        v = buf[role * 2048 + i]
        if v == 0:
            break
        ev.append((v >> 8, role, v & 255))
ev.sort()
t0 = ev[0][0]
for t, role, e in ev[: a.events]:
    print(f"{t - t0:8d}  {'   ' * role * 6}{['TMA', 'MMA', 'SOFTMAX'][role]}: {names[role][e]}")
