"""Where the MATH_F16F8 GEMM's warps wait: GNNLM_F8_DEBUG=32 makes gnnlm_linear_f16f8 print, per launch, the cycles the MMA issuer
spent waiting for a free TMEM accumulator / a full operand stage, the TMA producer for an empty stage, and the epilogue warps
waiting / working (clock64, averaged over the CTAs)."""
import os, subprocess, sys
if len(sys.argv) > 1:
    sys.path.insert(0, '.')
    sys.path.insert(0, 'profiles')
    from gemm_probe_f16f8 import run
    for shape in ((292040, 3072, 1024), (292040, 1024, 1024)):
        run(*shape, check=False)
else:
    for dbg in [int(x) for x in os.environ.get('DBGS', '32,40,36').split(',')]:
        env = dict(os.environ, GNNLM_F8_DEBUG=str(dbg))
        p = subprocess.run([sys.executable, __file__, "x"], env=env, capture_output=True, text=True)
        lines = [l for l in p.stderr.splitlines() if l.startswith('f16f8')]
        print(f"--- GNNLM_F8_DEBUG={dbg}\n{p.stdout}" + "\n".join(lines[-8:]), flush=True)
