# Round-1 evidence pass (after the tensor-core inter kernel / unrolled w = 5, 7 cluster kernels): run on the GPU box via gpurun.
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
# launch list of exactly one resident step (cudaProfilerStart/Stop range)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches.csv python bench.py --ncu-range --no-cpu-baseline --also-modes "" --n-datastore 16777216 > gpurun_out/final_launches.log 2>&1
# full counters of the HBM-bound kernels of that step, and of five projection GEMM launches
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"cluster_attn|edge_attn|inter_mma|pq_decode|layernorm|causal_softmax|knn_mix|gather_rows" -o gpurun_out/final_hbm python bench.py --ncu-range --no-cpu-baseline --also-modes "" --n-datastore 16777216 > gpurun_out/final_hbm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_f16s" -c 5 -o gpurun_out/final_gemm python bench.py --ncu-range --no-cpu-baseline --also-modes "" --n-datastore 16777216 > gpurun_out/final_gemm.log 2>&1
# other shapes
timeout 300 python bench.py --config c1 --no-cpu-baseline --also-modes "" > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
timeout 300 python bench.py --config c2 --no-cpu-baseline --also-modes "" > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 600 python bench.py --config c4 --no-cpu-baseline --also-modes "" > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 300 python bench.py --deprecated --no-cpu-baseline --also-modes "" > gpurun_out/bench_dedup.json 2> gpurun_out/bench_dedup.err
# BASELINE.json configs[4]: k x c sweep
SWEEP_OUT=gpurun_out/r1_sweep_c5.json timeout 1500 python profiles/sweep_c5.py > gpurun_out/sweep.log 2>&1
tail -3 gpurun_out/sweep.log | cut -c1-300
ls -la gpurun_out/ | head -40
