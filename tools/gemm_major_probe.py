"""dgrad / wgrad GEMM shapes of the training step: K-major operands (explicitly transposed copies) against the MN-major
operand modes that read the tensors as stored.  CUDA events, L2 flushed between launches."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_latent_diffusion_b200 import _lib
L = _lib.load()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn):
    ts = []
    for i in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

T = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
print(f"tokens T = {T}")
for name, (M, N, K) in {"dgrad d_hid2 = dy W_down": (T, 3072, 768), "dgrad dxn = d_hid W_up": (T, 768, 3072),
                        "dgrad dxn = dq W_q": (T, 768, 768), "dgrad dxn = dqkv W_qkv": (T, 768, 2304)}.items():
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = torch.randn(K, N, device="cuda").bfloat16()          # stored [K = out features, N = in features]
    Wt = W.t().contiguous()                                  # [N, K]: what the K-major path needs
    out = torch.empty(M, N, device="cuda")
    t_k = timeit(lambda: _lib.check(L.tld_op_gemm(4, _lib.ptr(A), _lib.ptr(Wt), M, N, K, _lib.ptr(out), None, st), "k"))
    t_n = timeit(lambda: _lib.check(L.tld_op_gemm_nn(4, _lib.ptr(A), _lib.ptr(W), M, N, K, _lib.ptr(out), st), "n"))
    fl = 2.0 * M * N * K
    print(f"{name:28s} M={M} N={N} K={K}:  K-major {t_k:7.1f} us ({fl/t_k/1e6:6.0f} TF/s)   MN-major B {t_n:7.1f} us ({fl/t_n/1e6:6.0f} TF/s)")
for name, (M, N, K) in {"wgrad dW_down = dy^T hid2": (768, 3072, T), "wgrad dW_up = d_hid^T xn": (3072, 768, T),
                        "wgrad dW_qkv = dqkv^T xn": (2304, 768, T)}.items():
    A = torch.randn(K, M, device="cuda").bfloat16()
    B = torch.randn(K, N, device="cuda").bfloat16()
    At, Bt = A.t().contiguous(), B.t().contiguous()
    out = torch.empty(M, N, device="cuda")
    t_k = timeit(lambda: _lib.check(L.tld_op_gemm(4, _lib.ptr(At), _lib.ptr(Bt), M, N, K, _lib.ptr(out), None, st), "k"))
    t_n = timeit(lambda: _lib.check(L.tld_op_gemm_mn(4, _lib.ptr(A), _lib.ptr(B), M, N, K, _lib.ptr(out), st), "n"))
    fl = 2.0 * M * N * K
    print(f"{name:28s} M={M} N={N} K={K}:  K-major {t_k:7.1f} us ({fl/t_k/1e6:6.0f} TF/s)   MN-major A,B {t_n:7.1f} us ({fl/t_n/1e6:6.0f} TF/s)")
