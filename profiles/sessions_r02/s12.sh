#!/bin/bash
# GPU session (1 GPU): new multigrid cycle variants + streaming Gram-Schmidt + C4 harness tests; C5 long run (20 steps) with
# V(1,1), and V(0,1) / V(1,0) / V(0,2) for comparison; C4 adaptive run on one GPU
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zmultigrid.py tests/test_gpu_parity.py tests/test_gpu_driver.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/s12_tests.log 2>&1
tail -4 gpurun_out/s12_tests.log
B="python bench.py --no-c2 --no-cpu-baseline"
timeout 300 $B --steps 20 --warmup 3 > gpurun_out/s12_c5_v11_k20.json 2> gpurun_out/s12_c5_v11_k20.err
for v in "0 1" "1 0" "0 2"; do
  set -- $v
  timeout 300 $B --steps 10 --warmup 3 --mg-pre $1 --mg-post $2 > gpurun_out/s12_c5_v$1$2.json 2> gpurun_out/s12_c5_v$1$2.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/s12_c5_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.2f" % j["ms_per_step"], j["preconditioner"][:24], j["newton"]["first_run_gmres_its"], "steps_to_converge", j["newton"]["steps_to_converge"],
              "inner %.2f" % j["ms_per_inner_step"], "bj", j.get("block_jacobi", {}).get("ms_per_step"), j.get("preconditioner_fallback"))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 300 python tools/c4_adaptive.py --json gpurun_out/s12_c4_1gpu.json > gpurun_out/s12_c4_1gpu.log 2>&1
tail -8 gpurun_out/s12_c4_1gpu.log
