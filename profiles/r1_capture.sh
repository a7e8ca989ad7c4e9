set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --also-modes "" --no-cuda-graph --n-datastore 16777216 > gpurun_out/final_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"cluster_attn|edge_attn|pq_decode|layernorm|causal_softmax|knn_mix" -o gpurun_out/final_hbm python bench.py --ncu-range --no-cpu-baseline --also-modes "" --n-datastore 16777216 > gpurun_out/final_hbm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_f16s" -c 5 -o gpurun_out/final_gemm python bench.py --ncu-range --no-cpu-baseline --also-modes "" --n-datastore 16777216 > gpurun_out/final_gemm.log 2>&1
ls -la gpurun_out/
