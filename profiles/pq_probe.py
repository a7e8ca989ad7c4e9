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
