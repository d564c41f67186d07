#!/bin/bash
# GPU session (2 GPUs): shared communicator (vh_comm_share) in the multigrid hierarchy and across adaptive cycles
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_driver.py -m gpu -q --tb=short -p no:cacheprovider -k "c4 or two_gpu" > gpurun_out/s18_tests.log 2>&1
tail -4 gpurun_out/s18_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29811 tools/c4_adaptive.py --cycles 3 --initial-refine 3 --json gpurun_out/s18_c4_share.json > gpurun_out/s18_c4_share.log 2>&1
timeout 200 $TR --master-port 29812 tools/c4_adaptive.py --cycles 3 --initial-refine 3 --new-comm-per-cycle --json gpurun_out/s18_c4_newcomm.json > gpurun_out/s18_c4_newcomm.log 2>&1
python - <<'PY'
import json
for f in ("share", "newcomm"):
    try:
        j = json.load(open("gpurun_out/s18_c4_%s.json" % f))
        print(f, "t_context_s per cycle", ["%.2f" % c["t_context_s"] for c in j["cycles"]], "its", [h["linear_its"] for h in j["history"]])
    except Exception as e:
        print(f, "ERR", e)
PY
