"""ntgt -> tgt (inter) attention kernel alone at the Wiki103 shape: time per launch and algorithmic GB/s.
GNNLM_INTER_WARPS=8 selects the 8-warp form (default 16)."""
import os, sys, torch
sys.path.insert(0, '.')
from gnnlm_b200 import ops, _lib as L
dev = torch.device('cuda')
T, k, d, H = 3072, 32, 1024, 8
g = torch.Generator(device=dev).manual_seed(0)
deg = torch.full((T,), k, dtype=torch.int32, device=dev)
deg[torch.rand(T, generator=g, device=dev) < 0.3] -= 1
indptr = torch.zeros(T + 1, dtype=torch.int32, device=dev)
indptr[1:] = torch.cumsum(deg, 0)
n_c = int(indptr[-1])
hc = ops.to_split(torch.randn((n_c, d), generator=g, device=dev))
q = torch.randn((T, d), generator=g, device=dev)
Wk = torch.randn((d, d), generator=g, device=dev) / 32
Wv = torch.randn((d, d), generator=g, device=dev) / 32
bv = torch.randn(d, generator=g, device=dev)
dk = d // H
wk_t = ops.split_f16(Wk.view(H, dk, d).transpose(1, 2).contiguous().view(H * d, dk))
wv = ops.split_f16(Wv.contiguous())
out = torch.empty((T, d), device=dev)
f = lambda: ops.inter_attn_fused(q, [(0, T, indptr, hc)], H, out, wk_t, wv, bv, out_scale=0.5)
for _ in range(3):
    f()
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L.TIMING = []
for _ in range(10):
    big.zero_()                       # flush L2 between launches
    f()
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for n, tag, a, b, w in L.TIMING if tag == "inter_fused"]
L.TIMING = None
ms = sorted(ms)[len(ms) // 2]
byt = n_c * d * 4 + 2 * T * H * d * 4 + T * d * 4
print(f"inter kernel (warps={os.environ.get('GNNLM_INTER_WARPS', '16')}): {ms * 1e3:.1f} us, {byt / ms / 1e6:.0f} GB/s algorithmic ({byt / 1e6:.0f} MB)")
