"""f16x3 GEMM with split-fp16 activations: which of (A split, C split, residual split) costs what."""
import sys, torch
sys.path.insert(0, '.')
from gnnlm_b200 import ops, _lib as L
dev = torch.device('cuda')
def run(M, N, K, a_split, c_split, res):   # res in (None, "f32", "split")
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / 32
    Wh, Wl, sc = ops.split_f16(W)
    Ain = ops.to_split(A) if a_split else A
    R = None
    if res: R = torch.randn(M, N, device=dev); R = ops.to_split(R) if res == "split" else R
    out = ops.empty_act(M, N, ops.SPLIT if c_split else torch.float32, dev)
    f = lambda: ops.linear(Ain, Wh, None, W_lo=Wl, w_scale=sc, residual=R, out=out, math=L.MATH_F16X3)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"M={M} N={N} K={K} A_split={a_split} C_split={c_split} res={res}: {ms:.3f} ms {2*M*N*K/ms/1e9:.0f} TF/s")
for N in (1024, 3072):
    run(292040, N, 1024, False, False, None)
    run(292040, N, 1024, True, False, None)
    run(292040, N, 1024, True, True, None)
    run(292040, N, 1024, True, False, "f32")
    run(292040, N, 1024, True, False, "split")
