#!/bin/bash
# GPU session (8 GPUs): default bench at N=8 (C5, multigrid V(0,1), block-Jacobi next to it, parity against 1 GPU), the multigrid
# parity worker on 8 ranks, BASELINE configs[3] (adaptive slab) on 8 ranks with the P-independence check
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time timeout 150 $TR --master-port 29801 bench.py --gpus 8 > gpurun_out/s17_bench_n8.json 2> gpurun_out/s17_bench_n8.err ) 2> gpurun_out/s17_bench_n8.time
tail -3 gpurun_out/s17_bench_n8.time
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/s17_bench_n8.json").read().strip().splitlines()[-1])
    print("N=8 ms/step %.2f" % j["ms_per_step"], j["preconditioner"][:24], j["newton"]["first_run_gmres_its"], "inner %.2f" % j["ms_per_inner_step"], j["phase_ms_per_step"],
          "bj", j.get("block_jacobi", {}).get("ms_per_step"), "halo", j.get("halo_ms_per_exchange"), "parity", j.get("multi_gpu_parity"), j.get("preconditioner_fallback"))
except Exception as e:
    print("bench N=8 ERR", e)
PY
timeout 80 $TR --master-port 29802 tests/multigpu_worker.py mg > gpurun_out/s17_mg_worker_n8.log 2>&1; grep -E "PARITY|Error" gpurun_out/s17_mg_worker_n8.log | tail -2
( time timeout 220 $TR --master-port 29803 tools/c4_adaptive.py --check-single --json gpurun_out/s17_c4_8gpu.json > gpurun_out/s17_c4_8gpu.log 2>&1 ) 2> gpurun_out/s17_c4_8gpu.time
grep -E "^cycle|P-INDEP|DONE|Error" gpurun_out/s17_c4_8gpu.log | cut -c1-700 | tail -10
tail -3 gpurun_out/s17_c4_8gpu.time
