"""Pure-write and copy bandwidth on this GPU (what bounds the streaming-write kernels, e.g. PQ decode: 1.2 GB written, 37 MB read)."""
import torch
dev = torch.device("cuda")
x = torch.empty(1 << 30, dtype=torch.float32, device=dev)          # 4 GiB
y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms = t(lambda: x.zero_())
print(f"memset 4 GiB: {ms:.3f} ms -> {x.numel() * 4 / ms / 1e9:.2f} TB/s written")
ms = t(lambda: x.fill_(1.5))
print(f"fill   4 GiB: {ms:.3f} ms -> {x.numel() * 4 / ms / 1e9:.2f} TB/s written")
ms = t(lambda: y.copy_(x))
print(f"copy   4 GiB: {ms:.3f} ms -> {2 * x.numel() * 4 / ms / 1e9:.2f} TB/s read + written")
ms = t(lambda: x.sum())
print(f"read   4 GiB: {ms:.3f} ms -> {x.numel() * 4 / ms / 1e9:.2f} TB/s read")
