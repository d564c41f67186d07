#!/bin/bash
# GPU session (2 GPUs): the whole GPU test-suite (2-GPU tests included), default bench at N=2, C4 adaptive on 2 GPUs with the 1-GPU check
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/s13_gpu_tests_2gpu.log 2>&1
tail -6 gpurun_out/s13_gpu_tests_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time timeout 600 $TR --master-port 29711 bench.py --gpus 2 > gpurun_out/s13_bench_n2.json 2> gpurun_out/s13_bench_n2.err ) 2> gpurun_out/s13_bench_n2.time
tail -3 gpurun_out/s13_bench_n2.time
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/s13_bench_n2.json").read().strip().splitlines()[-1])
    print("N=2 ms/step %.2f" % j["ms_per_step"], j["preconditioner"][:24], j["newton"]["first_run_gmres_its"], "inner %.2f" % j["ms_per_inner_step"],
          "bj", j.get("block_jacobi", {}).get("ms_per_step"), "parity", j.get("multi_gpu_parity"), j.get("preconditioner_fallback"))
except Exception as e:
    print("bench N=2 ERR", e)
PY
( time timeout 600 $TR --master-port 29712 tools/c4_adaptive.py --check-single --json gpurun_out/s13_c4_2gpu.json > gpurun_out/s13_c4_2gpu.log 2>&1 ) 2> gpurun_out/s13_c4_2gpu.time
grep -E "^cycle|P-INDEP|DONE|Error|error" gpurun_out/s13_c4_2gpu.log | tail -12
tail -3 gpurun_out/s13_c4_2gpu.time
