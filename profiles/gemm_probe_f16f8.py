"""MATH_F16F8 (fp16 main product + FP8 correction MMAs) against the 3xFP16 form: accuracy vs fp64 and time per launch."""
import sys, torch
sys.path.insert(0, '.')
from gnnlm_b200 import ops, _lib as L
dev = torch.device('cuda')


def timeit(f, n=5):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(M, N, K, K2=0, check=True):
    g = torch.Generator(device=dev).manual_seed(M + N + K)
    A = torch.randn((M, K), generator=g, device=dev)
    A2 = torch.randn((M, K2), generator=g, device=dev) if K2 else None
    W = torch.randn((N, K + K2), generator=g, device=dev) / 32
    b = torch.randn((N,), generator=g, device=dev)
    Wh, Wl, sc = ops.split_f16(W)
    W8 = ops.quant_w8(Wh, Wl)
    As = ops.to_q8(ops.to_split(A))
    A2s = ops.to_q8(ops.to_split(A2)) if K2 else None
    out8 = torch.empty((M, N), device=dev)
    f8 = lambda: ops.linear_f16f8(As, Wh, W8, b, A2=A2s, w_scale=sc, out=out8)
    ms8 = timeit(f8)
    line = f"M={M} N={N} K={K}+{K2}: f16f8 {ms8:.3f} ms ({2 * M * N * (K + K2) / ms8 / 1e9:.0f} TF/s useful)"
    if not K2:
        out3 = torch.empty((M, N), device=dev)
        f3 = lambda: ops.linear(As, Wh, b, W_lo=Wl, w_scale=sc, out=out3, math=L.MATH_F16X3)
        ms3 = timeit(f3)
        line += f", f16x3 {ms3:.3f} ms, ratio {ms3 / ms8:.2f}"
    if check:
        rows = torch.randint(0, M, (512,), device=dev)
        Af = A if not K2 else torch.cat([A, A2], 1)
        ref = Af[rows].double() @ W.double().T + b.double()
        scale = ref.abs().max().item()
        e8 = (out8[rows].double() - ref).abs().max().item() / scale
        line += f"; max err / max|C|: f16f8 {e8:.2e}"
        if not K2:
            line += f", f16x3 {(out3[rows].double() - ref).abs().max().item() / scale:.2e}"
        # single fp16 pass for scale: hi halves only
        one = As.data[rows, :K].double() @ (Wh.double() / sc).T[:K] + (0 if not K2 else A2s.data[rows, :K2].double() @ (Wh.double() / sc).T[K:]) + b.double()
        line += f", one fp16 pass {(one - ref).abs().max().item() / scale:.2e}"
    print(line, flush=True)


if __name__ == "__main__":
    run(1000, 300, 128)
    run(4096, 3072, 1024)
    run(292040, 3072, 1024)
    run(292040, 2048, 1024)
    run(292040, 1024, 1024)
    run(97400, 1024, 1024)
    run(292040, 1024, 1024, K2=1024)
    run(3072, 1024, 1024)
