"""Timing probe for the tcgen05 GEMM kernels (run on the GPU box): python profiles/gemm_probe.py [math]"""
import os, sys, torch
sys.path.insert(0, '.')
from gnnlm_b200 import ops, _lib as L
dev = torch.device('cuda')
math = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
mode = L.MATH_NAMES[math]
def run(M, N, K, residual=False, bias=False, m_live=None):
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / 32
    lo, sc = None, 1.0
    if math == "f16x3": W, lo, sc = ops.split_f16(W)
    elif math == "tf32x3": W, lo = ops.split_tf32(W)
    elif math == "bf16": A, W = A.bfloat16(), W.bfloat16()
    out = torch.empty(M, N, device=dev)
    R = torch.randn(M, N, device=dev) if residual else None
    b = torch.randn(N, device=dev) if bias else None
    md = torch.tensor([m_live], dtype=torch.int32, device=dev) if m_live else None
    f = lambda: ops.linear(A, W, b, W_lo=lo, w_scale=sc, residual=R, out=out, m_dev=md, math=mode)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    Ml = m_live or M
    print(f"{math} dbg={os.environ.get('GNNLM_GEMM_DEBUG','0')} M={M} live={Ml} N={N} K={K} res={residual} bias={bias}: {ms:.3f} ms  {2*Ml*N*K/ms/1e9:.1f} TFLOP/s")
run(292040, 3072, 1024, bias=True)
run(292040, 2048, 1024, bias=True)
run(292040, 1024, 1024)
run(292040, 1024, 1024, residual=True, bias=True)
run(294912, 1024, 1024, residual=True, bias=True, m_live=292040)
run(98304, 1024, 1024, residual=True, bias=True, m_live=97382)
run(98304, 2048, 1024, bias=True, m_live=97382)
run(3072, 3072, 1024, bias=True)
