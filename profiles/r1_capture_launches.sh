# launch list of exactly one resident step (cudaProfilerStart/Stop range), then the C5 sweep
cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches.csv python bench.py --ncu-range --no-cpu-baseline --also-modes "" --n-datastore 16777216 > gpurun_out/final_launches.log 2>&1
SWEEP_OUT=gpurun_out/r1_sweep_c5.json timeout 1500 python profiles/sweep_c5.py > gpurun_out/sweep.log 2>&1
tail -3 gpurun_out/sweep.log | cut -c1-300
