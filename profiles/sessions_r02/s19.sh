#!/bin/bash
# final GPU session of the round (1 GPU): whole GPU test-suite, default bench line, DRAM traffic of the apply kernels at C5, launch list
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/s19_gpu_tests.log 2>&1
tail -4 gpurun_out/s19_gpu_tests.log
( time timeout 200 python bench.py > gpurun_out/s19_bench_default.json 2> gpurun_out/s19_bench_default.err ) 2> gpurun_out/s19_bench_default.time
tail -3 gpurun_out/s19_bench_default.time
B="python bench.py --steps 2 --warmup 1 --no-c2 --no-cpu-baseline"
timeout 150 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k 'regex:k_points|k_gather_apply' -c 5 --csv --log-file gpurun_out/s19_traffic_c5.csv $B > gpurun_out/s19_ncu_traffic.log 2>&1
grep -c k_points gpurun_out/s19_traffic_c5.csv
timeout 330 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/s19_launches_default.csv $B \
  > gpurun_out/s19_ncu_launches.log 2>&1
wc -l gpurun_out/s19_launches_default.csv
head -c 600 gpurun_out/s19_bench_default.json
