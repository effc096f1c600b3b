import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_latent_diffusion_b200 import _lib
B, n_tok, D = [int(v) for v in sys.argv[1:4]]
impl = int(sys.argv[4]) if len(sys.argv) > 4 else 3
L = _lib.load()
g = torch.Generator(device="cuda").manual_seed(5)
T = B * n_tok
qkv = torch.randn(T, 3 * D, device="cuda", generator=g).bfloat16()
H = D // 64
q, k, v = (t.float().view(B, n_tok, H, 64).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=1))
o = (torch.softmax((q @ k.transpose(-1, -2)) / 8.0, -1) @ v)  # B,H,N,64
for rep in range(3):
    x = torch.zeros(T, D, device="cuda")
    _lib.check(L.tld_op_self_attention(_lib.ptr(qkv), _lib.ptr(x), B, n_tok, D, impl, torch.cuda.current_stream().cuda_stream), "a")
    torch.cuda.synchronize()
    got = x.view(B, n_tok, H, 64).permute(0, 2, 1, 3)
    err = (got - o).view(B, H, n_tok // 128, 128, 64)
    rel = err.flatten(3).norm(dim=3) / o.view(B, H, n_tok // 128, 128 * 64).norm(dim=3)
    bad = (rel > 0.02).nonzero().tolist()
    print(f"rep {rep}: {len(bad)} bad tiles of {rel.numel()}; first: ", [(b, h, t, (b * H + h) * (n_tok // 128) + t, round(rel[b, h, t].item(), 3)) for b, h, t in bad[:12]])
    if bad:
        b, h, t = bad[0]
        e = err[b, h, t]  # 128 x 64
        print("  row-block errors (4 x 32 rows) x (2 x 32 cols):", [[round(e[i*32:(i+1)*32, j*32:(j+1)*32].norm().item(), 3) for j in range(2)] for i in range(4)])
        print("  got/ref ratio sample:", (got[b, h, t*128:(t*128+4), :4] / o[b, h, t*128:(t*128+4), :4]).tolist())
