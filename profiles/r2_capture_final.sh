#!/bin/bash
# Round-2 final evidence on one B200: the default bench line, the reference arm, the launch list of exactly one resident step,
# ncu --set full of the tcgen05 flash kernel, the role-wait logs of the two tcgen05 kernels.
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_reference.json 2> gpurun_out/r2_final_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_final_launches.csv \
  python bench.py --ncu-range --no-cpu-baseline --also-modes "" --locality "" --n-datastore 16777216 > gpurun_out/r2_final_launches.log 2>&1
FLASH_TC=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:causal_flash_tc -c 1 -s 3 -o gpurun_out/r2_flash_tc \
  python profiles/flash_probe.py x > gpurun_out/ncu_flash_tc.log 2>&1
FLASH_TC=1 python profiles/flash_probe.py > gpurun_out/r2_flash_tc_probe.log 2>&1
python profiles/flash_probe.py >> gpurun_out/r2_flash_tc_probe.log 2>&1
GNNLM_FLASH_DEBUG=1 FLASH_TC=1 python profiles/flash_probe.py x 2>&1 | tail -2 >> gpurun_out/r2_flash_tc_probe.log
tail -c 600 gpurun_out/r2_final_bench.err; cat gpurun_out/r2_flash_tc_probe.log
