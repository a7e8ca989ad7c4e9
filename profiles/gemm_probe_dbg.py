"""f16x3 QKV-shape GEMM under GNNLM_GEMM_DEBUG bits (1: splitter off, 2: W loads off, 4: MMAs off)."""
import os, sys, torch
sys.path.insert(0, '.')
from gnnlm_b200 import ops, _lib as L
dev = torch.device('cuda')
M, N, K = 292040, 3072, 1024
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / 32
hi, lo, sc = ops.split_f16(W)
out = torch.empty(M, N, device=dev)
f = lambda: ops.linear(A, hi, None, W_lo=lo, w_scale=sc, out=out, math=L.MATH_F16X3)
for _ in range(2): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): f()
e1.record(); torch.cuda.synchronize()
print(f"dbg={os.environ.get('GNNLM_GEMM_DEBUG','0')} {e0.elapsed_time(e1)/5:.3f} ms")
