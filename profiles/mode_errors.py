"""Whole-path log-prob error of every arithmetic mode against the fp32 CPU oracle (c3mini: d=1024, 3 HGT layers,
adaptive softmax).  Prints max / mean relative error of per-token log-probs and the NLL shift."""
import sys, copy, numpy as np, torch
sys.path.insert(0, '.')
from tests.synth import make_problem, run_oracle, run_gpu
prob = make_problem(sys.argv[1] if len(sys.argv) > 1 else "c3mini")
ref = run_oracle(prob)
r64 = run_oracle(prob, dtype=torch.float64)
lp64 = r64["logprob"].numpy()
print(f"oracle fp32 vs fp64: max rel {np.abs(ref['logprob'].numpy()-lp64).max()/np.abs(lp64).mean():.2e}")
for mode in ("fp32", "tf32x3", "f16x3", "tf32", "bf16"):
    out = run_gpu(prob, torch.device("cuda"), math=mode)
    err = np.abs(out["logprob"] - lp64)
    rel = err / np.abs(lp64)
    print(f"{mode:7s} max rel {rel.max():.2e}  mean rel {rel.mean():.2e}  max abs {err.max():.2e}  d nll {abs(out['nll'] - r64['nll']):.2e}")
