"""PQ gather + decode on its own (wiki103 shape: 292k nodes, M=128, dsub=8): GB/s of the fp32-codebook and pre-split paths."""
import json, os, sys, torch
sys.path.insert(0, '.')
from gnnlm_b200 import ops, synth
dev = torch.device('cuda')
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if os.path.exists('MEASURED_PEAKS.json') else 6650.0
cfg = dict(synth.CONFIGS['c3'], n_d=1 << 24)
model = synth.make_model(cfg)
q = model.decoder.tgt_quantizer.to(dev)
tables = synth.make_tables(cfg, device=dev)
n = 292040
rows = torch.randint(0, tables['n_d'], (n,), device=dev)
hi, lo = q._split_codebook()
by = n * (128 + 8 + 4096)
def t(f, reps=20):
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, f in (("fp32 codebook -> split", lambda: ops.pq_gather_decode(tables['codes'], q.centroids_torch, rows, out_dtype=ops.SPLIT)),
                ("fp32 codebook -> fp32", lambda: ops.pq_gather_decode(tables['codes'], q.centroids_torch, rows, out_dtype=torch.float32)),
                ("pre-split codebook", lambda: ops.pq_gather_decode_presplit(tables['codes'], hi, lo, rows))):
    ms = t(f)
    print(f"{name:26s} {ms:.3f} ms  {by / ms / 1e6:.0f} GB/s ({by / ms / 1e6 / peak * 100:.0f}%)", flush=True)
q8cb = q._q8_codebook()
ms = t(lambda: ops.pq_gather_decode_hiq8(tables['codes'], hi, q8cb, rows))
print(f"{'hi + e4m3 companion (default)':26s} {ms:.3f} ms  {by / ms / 1e6:.0f} GB/s ({by / ms / 1e6 / peak * 100:.0f}%)", flush=True)
# two-phase form: gather the code rows compactly first (one coalesced 128 B row per node), then decode from sequential codes
ar = torch.arange(n, device=dev)
gather = lambda: ops.pq_gather_decode(tables['codes'], q.centroids_torch, rows, want_codes=True, decode=False)[2]
ms_g = t(gather)
compact = gather()
ms_d = t(lambda: ops.pq_gather_decode_hiq8(compact, hi, q8cb, ar))
print(f"two-phase: gather codes {ms_g:.3f} ms + decode from compact codes {ms_d:.3f} ms = {ms_g + ms_d:.3f} ms", flush=True)
