"""Where evaluate()'s wall time goes at the Wiki103 shape: producer threads vs device step."""
import os, sys, time, shutil
sys.path.insert(0, '.')
import torch
import bench
from gnnlm_b200 import synth
from gnnlm_b200.eval_lm import evaluate
dev = torch.device('cuda')
print("host cores:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))
cfg = dict(synth.CONFIGS["c3"], n_d=1 << 24)
model = synth.make_model(cfg).to(dev).set_math("f16f8")
root = bench.tmp_root("probe")
try:
    ds, dstore, knn, scorer, info = bench.build_eval(cfg, model, dev, "f16f8", root, 48, tail=0)
    cache = {}
    def run(**kw):
        kw = dict(kw, graph_cache=cache)
        evaluate(model, ds, dstore, scorer, knn_dstore=knn, temperature=1.0, max_sentences=1, device=dev, cuda_graph=True, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = evaluate(model, ds, dstore, scorer, knn_dstore=knn, temperature=1.0, max_sentences=1, device=dev, cuda_graph=True, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        kw.pop("graph_cache")
        print(kw, f"{r['count'] / dt:.0f} tokens/s, {dt / len(ds) * 1e3:.2f} ms per block",
              {k: round(v * 1e3 / len(ds), 2) for k, v in r["host_profile"].items()}, flush=True)
    for w in (0, 0, 1, 2):
        run(host_workers=w)
    for t in (1, 8):
        ds.host_copy_threads = t
        run(host_workers=0)
    run(host_threads=False)
    # producer alone
    t0 = time.perf_counter()
    spec = ds.batch_spec([0]); bufs = {n: torch.empty(sh, dtype=dt, pin_memory=True) for n, sh, dt in spec}
    for i in range(len(ds)):
        ds.collate_into([i], bufs)
    print(f"collate_into alone: {(time.perf_counter() - t0) / len(ds) * 1e3:.2f} ms per block")
finally:
    shutil.rmtree(root, ignore_errors=True)
