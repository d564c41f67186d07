#!/bin/bash
# One standard GPU session (run under gpurun from the repo root).  Everything lands in gpurun_out/; turn the raw files into
# the committed summaries with profiles/summarize.py (see its header).
#   gpurun --timeout 900 -- 'bash tools/gpu_session.sh [tests|bench|ncu|all]'          (one GPU)
#   gpurun --gpus N --timeout 600 -- 'bash tools/gpu_multi.sh N'                        (multi-rank parity + bench; never under ncu)
# ncu replays every profiled launch: the captures below use a SHORT bench command (2 steps) and metric subsets.
set -u
what=${1:-all}
mkdir -p gpurun_out
if [ "$what" = tests ] || [ "$what" = all ]; then
  timeout 700 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
  tail -5 gpurun_out/gpu_tests.log
fi
if [ "$what" = bench ] || [ "$what" = all ]; then
  ( time timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time
  tail -3 gpurun_out/bench_default.time
  python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
    print("ms/step %.2f" % j["ms_per_step"], j["preconditioner"], j["newton"]["first_run_gmres_its"], "e2e %.2f" % j["e2e"]["ms_per_step"],
          j["phase_ms_per_step"], "bj", j.get("block_jacobi", {}).get("ms_per_step"), "roofline frac %.3f" % j["roofline"]["frac"],
          "c2 %.3f" % j["c2"]["ms_per_step"] if "c2" in j and "ms_per_step" in j["c2"] else "")
except Exception as e:
    print("bench ERR", e)
PY
fi
if [ "$what" = ncu ] || [ "$what" = all ]; then
  B="python bench.py --steps 2 --warmup 1 --no-c2 --no-cpu-baseline"
  # DRAM bytes of the dominant kernels of the headline workload: after the fine / level-6 / level-5 assembly launches of
  # k_points the next k_points launch is the fine-level operator apply, followed by its k_gather_apply
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k 'regex:k_points|k_gather_apply' -c 5 --csv --log-file gpurun_out/traffic_c5.csv $B > gpurun_out/ncu_traffic.log 2>&1
  # every launch of the bench command with its device time (shares of the step)
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_default.csv $B \
    > gpurun_out/ncu_launches.log 2>&1
  ls -la gpurun_out/traffic_c5.csv gpurun_out/launches_default.csv 2>/dev/null
fi
