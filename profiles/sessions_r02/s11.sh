#!/bin/bash
# GPU session (1 GPU): full GPU tests, the default bench line with its wall time, a traced short run, ncu of the block inverse
set -u
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/s11_gpu_tests.log 2>&1
tail -4 gpurun_out/s11_gpu_tests.log
( time timeout 600 python bench.py > gpurun_out/s11_bench_default.json 2> gpurun_out/s11_bench_default.err ) 2> gpurun_out/s11_bench_default.time
tail -3 gpurun_out/s11_bench_default.time
VH_GMRES_TRACE=1 VH_MG_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-c2 --no-cpu-baseline > gpurun_out/s11_trace.json 2> gpurun_out/s11_trace.err
grep -c "mg trace" gpurun_out/s11_trace.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_block_invert -c 3 -o gpurun_out/s11_inv -f python bench.py --steps 1 --warmup 1 --no-c2 --no-cpu-baseline > gpurun_out/s11_ncu_inv.log 2>&1
ls -la gpurun_out/s11_*
head -c 1500 gpurun_out/s11_bench_default.json
