#!/bin/bash
# Final round-2 evidence on one B200 (after the heterograph / training work): full GPU suite, the default bench line, the reference arm,
# the launch list of exactly one resident step, the fine-tuning step probe in both tensor-core modes.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/r2_gputest_full.log
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_reference.json 2> gpurun_out/r2_final_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_final_launches.csv \
  python bench.py --ncu-range --no-cpu-baseline --also-modes "" --locality "" --n-datastore 16777216 > gpurun_out/r2_final_launches.log 2>&1
python profiles/train_probe.py c3 f16x3 > gpurun_out/r2c_train_probe_f16x3.log 2>&1
python profiles/train_probe.py c3 tf32x3 > gpurun_out/r2c_train_probe_tf32x3.log 2>&1
tail -3 gpurun_out/r2_gputest_full.log; tail -c 400 gpurun_out/r2_final_bench.err; head -c 300 gpurun_out/r2_final_bench.json
