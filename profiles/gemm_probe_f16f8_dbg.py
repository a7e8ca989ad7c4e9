"""Where the MATH_F16F8 GEMM's time goes: GNNLM_F8_DEBUG bit 0 drops the FP8 products, bit 1 the fp16 product, bit 2 the epilogue,
bit 3 (8) only the epilogue's global stores, bit 4 (16) only its TMEM loads / shared-memory staging."""
import os, subprocess, sys
if len(sys.argv) > 1:
    sys.path.insert(0, '.')
    import torch
    from profiles.gemm_probe_f16f8 import run
    for shape in ((292040, 3072, 1024), (292040, 1024, 1024)):
        run(*shape, check=False)
else:
    for dbg in [int(x) for x in os.environ.get('DBGS', '0,1,2,3,4,7,8,16').split(',')]:
        env = dict(os.environ, GNNLM_F8_DEBUG=str(dbg))
        out = subprocess.run([sys.executable, __file__, "x"], env=env, capture_output=True, text=True).stdout
        print(f"--- GNNLM_F8_DEBUG={dbg}\n{out}", flush=True)
