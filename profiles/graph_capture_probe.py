"""Does one whole hot-path step capture into a CUDA graph (no host syncs, device-side counts)?  Times eager vs replay."""
import sys, torch
sys.path.insert(0, '.')
from gnnlm_b200 import synth
dev = torch.device('cuda')
for name in ("c1", "c2", "c3"):
    cfg = dict(synth.CONFIGS[name])
    if name == "c3": cfg["n_d"] = 1 << 24
    model = synth.make_model(cfg)
    tables = synth.make_tables(cfg, device=dev)
    batch = synth.make_batch(cfg, tables, device=dev)
    r = synth.Runner(cfg, model, tables, dev, "f16x3")
    for _ in range(3): r.step_resident(batch)
    torch.cuda.synchronize()
    def timeit(f, n=20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t_eager = timeit(lambda: r.step_resident(batch))
    acc0 = r.acc.clone()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        r.step_resident(batch)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        lp, _, _, _ = r.step_resident(batch)
    r.acc.zero_(); g.replay(); torch.cuda.synchronize()
    a1 = r.acc.clone(); r.acc.zero_(); r.step_resident(batch); torch.cuda.synchronize()
    same = torch.allclose(a1, r.acc, rtol=1e-12)
    t_graph = timeit(g.replay)
    T = cfg["B"] * cfg["L"]
    print(f"{name}: eager {t_eager:.3f} ms ({T/t_eager*1e3:.0f} tok/s)  graph replay {t_graph:.3f} ms ({T/t_graph*1e3:.0f} tok/s)  same result: {same}")
