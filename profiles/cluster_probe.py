"""ntgt-intra-ntgt cluster attention on its own: GB/s of the all-nodes and centre-only kernels per cluster size,
timed back to back (isolated: higher clocks than inside the power-capped step) -- for A/B-ing kernel changes.
  python profiles/cluster_probe.py [k] [reps]"""
import json, os, sys, torch
sys.path.insert(0, '.')
from gnnlm_b200 import ops, synth
from gnnlm_b200.graph import build_token_graph

dev = torch.device('cuda')
k = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if os.path.exists('MEASURED_PEAKS.json') else 6650.0
cfg = dict(synth.CONFIGS['c3'], n_d=1 << 22, k=k)
d, H, T = cfg['d'], cfg['H'], cfg['L']
tables = synth.make_tables(cfg, device=dev)
for c in (0, 1, 2, 3):
    batch = synth.make_batch(dict(cfg, c=c), tables, device=dev)
    g = build_token_graph(batch['nbr'], tables['n_d'], c, c)
    n, nv = g.counts()
    qkv = torch.randn(g.node_cap, 3 * d, device=dev)
    q, kk, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    qc = torch.randn(nv, d, device=dev)
    for out_kind in ('f32', 'split'):
        out = ops.empty_act(g.node_cap, d, ops.SPLIT if out_kind == 'split' else torch.float32, dev)
        outc = ops.empty_act(nv, d, ops.SPLIT if out_kind == 'split' else torch.float32, dev)
        res = {}
        for name, fn, by in (('full', lambda: ops.cluster_attn(q, kk, v, g, H, out), n * 4 * d * 4),
                             ('centre', lambda: ops.cluster_attn(qc, kk, v, g, H, outc, centre_only=True),
                              nv * (2 * min(2 * c + 1, 3) + 2) * d * 4)):
            for _ in range(3): fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            res[name] = (ms, by / ms / 1e6)
        print(f"k={k} c={c} w={2*c+1} out={out_kind}: " + "  ".join(f"{nm} {ms:.3f} ms {gbs:.0f} GB/s ({gbs/peak*100:.0f}%)" for nm, (ms, gbs) in res.items()), flush=True)
    del qkv, out, outc, g
