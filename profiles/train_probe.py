"""One fine-tuning step (forward + backward, train.train_step_loss) at the Wiki103 training shape of the reference's stage 1
(hgt_lm_wiki103_reproduce.sh:21-35: 3072-token samples, --gcn-k 32, --neighbor-context 1, 3 layers, --freeze): ms per step."""
import copy, json, sys, time, torch
sys.path.insert(0, '.')
from gnnlm_b200 import synth, train
dev = torch.device('cuda')
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
mode = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
cfg = dict(synth.CONFIGS[name], n_d=1 << 24)
model = synth.make_model(cfg).to(dev)
for n_, p in model.named_parameters():
    p.requires_grad_("hgt" in n_)
dropout = len(sys.argv) > 3 and sys.argv[3] == "dropout"          # model.train() with the rates of transformer_lm_wiki103
if dropout:
    model.train()
    for layer in model.decoder.hgt_decoder.gcs:
        layer.drop.p, layer.attn_drop.p = 0.3, 0.1
    if model.decoder.adaptive_softmax is not None:
        model.decoder.adaptive_softmax.dropout = 0.2
tables = synth.make_tables(cfg, device=dev)
batch = synth.make_batch(cfg, tables, device=dev)
r = synth.Runner(cfg, model, tables, dev, "fp32", prune_unreachable=False)
sample = r.sample_from(batch["nbr"], batch["feats"], batch["target"], batch["knn_dists"], batch["knn_ids"])
out = {}
for it in range(3):
    model.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss = train.train_step_loss(model, sample, mode)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    loss.backward()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    out = {"config": name, "math": mode, "dropout": dropout, "tokens": cfg["B"] * cfg["L"], "forward_ms": (t1 - t0) * 1e3, "backward_ms": (t2 - t1) * 1e3,
           "tokens_per_s": cfg["B"] * cfg["L"] / (t2 - t0), "loss_per_token": float(loss.detach()) / (cfg["B"] * cfg["L"]),
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
print(json.dumps(out))
# per-call breakdown of one more step (CUDA events around every C-ABI call)
from gnnlm_b200 import _lib as L
L.TIMING = []
model.zero_grad(set_to_none=True)
loss = train.train_step_loss(model, sample, mode)
n_fwd = len(L.TIMING)
loss.backward()
torch.cuda.synchronize()
agg = {}
for i, (nm, tag, a, b, work) in enumerate(L.TIMING):
    key = ("fwd " if i < n_fwd else "bwd ") + nm.replace("gnnlm_", "") + (f":{tag}" if tag and not tag.startswith("linear[") else "")
    v = agg.setdefault(key, [0.0, 0])
    v[0] += a.elapsed_time(b); v[1] += 1
timing, L.TIMING = L.TIMING, None
for k_, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"{k_:45s} {v[0]:9.2f} ms  x{v[1]}")
# the projections by shape: useful TFLOP/s of each (M, N, K)
shapes = {}
for i, (nm, tag, a, b, work) in enumerate(timing):
    if nm == "gnnlm_linear" and work:
        v = shapes.setdefault(("fwd" if i < n_fwd else "bwd",) + tuple(work), [0.0, 0])
        v[0] += a.elapsed_time(b); v[1] += 1
for k_, v in sorted(shapes.items(), key=lambda kv: -kv[1][0])[:12]:
    M_, N_, K_ = k_[1:]
    print(f"linear {k_[0]} M={M_} N={N_} K={K_}: {v[0]:8.2f} ms x{v[1]}  {2 * M_ * N_ * K_ * v[1] / v[0] / 1e9:7.1f} TFLOP/s useful")
