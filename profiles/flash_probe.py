"""gnnlm_hgt_causal_flash alone at the Wiki103 shape (L = 3072, H = 8, d_k = 128): time per launch and TFLOP/s of executed mma.sync work."""
import sys, torch
sys.path.insert(0, '.')
from gnnlm_b200 import ops
dev = torch.device('cuda')
B, L, H, d = 1, 3072, 8, 1024
torch.manual_seed(0)
qkv = torch.randn(B * L, 3 * d, device=dev) * 0.3
kv = ops.to_split(qkv[:, d:])
out = torch.zeros(B * L, d, device=dev)
res = ops.Split.empty(B * L, d, dev)
import os
if os.environ.get("FLASH_TC") == "1":
    qk = ops.to_split(qkv[:, :2 * d])
    vt = torch.empty((2, B * H * (d // H), L), device=dev, dtype=torch.float16)
    from gnnlm_b200 import _lib as LL
    vb = qkv[:, 2 * d:]
    LL.call("gnnlm_heads_transpose_split_f16", LL.ptr(vb), vb.stride(0), L, H, d // H, LL.ptr(vt[0]), LL.ptr(vt[1]), LL.stream_ptr())
    f = lambda: LL.call("gnnlm_hgt_causal_flash_tc", LL.ptr(qk.data), qk.data.stride(0), d, LL.ptr(vt), L, B, L, 0, H, d // H, LL.ptr(out),
                        out.stride(0), LL.ptr(res.data), res.data.stride(0), res.d, 0.5, 1, LL.stream_ptr())
else:
    f = lambda: ops.causal_attn_flash(qkv[:, :d], kv, B, L, 0, H, out, out_scale=0.5, accumulate=True, out_split=res)
for _ in range(3):
    f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 1 if len(sys.argv) > 1 else 20
e0.record()
for _ in range(n):
    f()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
flop = 2 * 2 * (L * (L + 1) / 2) * (d // H) * H * 3          # two products, three passes
print(f"causal_flash L={L}: {ms:.4f} ms per launch, {flop / ms / 1e9:.0f} TFLOP/s executed (mma.sync peak measured: 557 at 1.85 GHz)")
