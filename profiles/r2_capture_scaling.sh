#!/bin/bash
# 8-GPU box: strong scaling through eval_lm.evaluate (N = 1, 2, 4, 8 on one fixed corpus) and the weak line at N = 8
mkdir -p gpurun_out
python bench.py --gpus 1 --scaling strong > gpurun_out/r2_strong_n1.json 2> gpurun_out/r2_strong_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) \
    bench.py --gpus $n --scaling strong > gpurun_out/r2_strong_n$n.json 2> gpurun_out/r2_strong_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 \
  bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --also-modes "" > gpurun_out/r2_weak_n8.json 2> gpurun_out/r2_weak_n8.err
for f in gpurun_out/r2_strong_n*.json gpurun_out/r2_weak_n8.json; do
  python - "$f" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        p = json.loads(l)
        print(sys.argv[1], p["n_gpus"], p["scaling"], round(p["value"]), round(p["ms_per_step"], 2), p.get("score_sum"), p.get("count"))
PY
done
