"""BASELINE.json configs[4]: edge-aggregation scaling sweep on the Wiki103 shape -- k in {8,32,128,512} neighbours x
neighbour context c in {0,1,3} (cluster size w = 2c+1), one 3072-token block per step, 1 GPU.
Writes profiles/r1_sweep_c5.json (tokens/s, ntgt-intra-ntgt attention GB/s vs the measured HBM peak)."""
import json, os, sys, torch
sys.path.insert(0, '.')
from gnnlm_b200 import synth, _lib as L

dev = torch.device('cuda')
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if os.path.exists('MEASURED_PEAKS.json') else 6650.0
base = dict(synth.CONFIGS['c3'], n_d=1 << 24)
model = synth.make_model(base)
tables = synth.make_tables(base, device=dev)
rows = []
KS = [int(x) for x in os.environ.get('SWEEP_K', '8,32,128,512').split(',')]
CS = [int(x) for x in os.environ.get('SWEEP_C', '0,1,3').split(',')]
PRUNE = os.environ.get('SWEEP_PRUNE', '1') == '1'      # drop context nodes that cannot reach a tgt node (reach = NL-1)
for k in KS:
    for c in CS:
        cfg = dict(base, k=k, c=c)
        T, d = cfg['L'], cfg['d']
        c_eff = min(c, cfg['NL'] - 1) if PRUNE else c
        n_cap = T * k * (2 * c_eff + 1)
        est_gb = n_cap * d * 4 * 11 / 1e9          # live activation buffers of the ntgt side (x, h0, qkv(3), t, o, h1, kv(2))
        rec = dict(k=k, c=c, w_built=2 * c_eff + 1, n_ntgt_cap=n_cap, est_activation_gb=round(est_gb, 1))
        rec['token_chunked'] = est_gb > 48      # decoder budget: ntgt side runs in token chunks above 48 GB
        try:
            batch = synth.make_batch(cfg, tables, seed=k * 10 + c, device=dev)
            r = synth.Runner(cfg, model, tables, dev, 'f16x3', prune_unreachable=PRUNE)
            for _ in range(2 if est_gb < 100 else 1): r.step_resident(batch)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n_it = 3 if est_gb < 100 else 1
            for _ in range(n_it): r.step_resident(batch)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n_it
            L.TIMING = []
            r.step_resident(batch); torch.cuda.synchronize()
            t_nn = sum(a.elapsed_time(b) for n, tag, a, b, _w in L.TIMING if tag == 'nn_full')
            t_nc = sum(a.elapsed_time(b) for n, tag, a, b, _w in L.TIMING if tag == 'nn_centre')
            L.TIMING = None
            g = synth.build_token_graph(batch['nbr'], tables['n_d'], c_eff, c_eff)
            n_ntgt, n_valid = g.counts()
            del g
            rec.update(status='ok', ms_per_step=ms, tokens_per_s=T / ms * 1e3, n_ntgt=n_ntgt, n_valid=n_valid)
            if t_nn > 0:
                by = n_ntgt * 3 * d * 4 + n_ntgt * d * 4
                rec.update(nn_full_ms=t_nn, nn_full_gbs=by / t_nn / 1e6, nn_full_frac=by / t_nn / 1e6 / peak)
            if t_nc > 0:
                by = n_valid * (2 * min(2 * c + 1, 3) + 2) * d * 4      # K',V' of centre and its <= 2 chain neighbours, Q, out
                rec.update(nn_centre_ms=t_nc, nn_centre_gbs=by / t_nc / 1e6, nn_centre_frac=by / t_nc / 1e6 / peak)
            del r, batch
        except torch.OutOfMemoryError as e:
            rec['status'] = 'OOM'
        torch.cuda.empty_cache()
        rows.append(rec); print(rec, flush=True)
json.dump(rows, open(os.environ.get('SWEEP_OUT', 'gpurun_out/r1_sweep_c5.json'), 'w'), indent=1)
