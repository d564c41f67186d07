#!/bin/bash
# Multi-rank GPU session (run under `gpurun --gpus N`): parity of the N-rank run with the 1-rank run on three mesh types, with
# the NCCL ghost refresh and with the fused ghost push, the streaming Gram-Schmidt variant, and the bench on N GPUs.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_multi.sh 2 [bench-args...]'
set -u
N=${1:-2}; shift || true
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29600
run_worker() { # $1 = mode, rest = env assignments
  local mode=$1; shift
  port=$((port + 1))
  local arg=$mode; [ "$mode" = cube ] && arg=""
  local tag="mg${N}_${mode}_$(echo "$*" | tr ' =' '__')"
  ( env "$@" timeout 300 $TR --master-port $port tests/multigpu_worker.py $arg 2>&1 | grep -E "PARITY|Error|error|hist" | tail -3 ) > gpurun_out/$tag.log
  echo "$mode [$*]: $(grep -E 'PARITY' gpurun_out/$tag.log | tail -1)"
}
for mode in cube periodic hanging; do
  run_worker $mode VH_HALO_PUSH=0
  run_worker $mode VH_HALO_PUSH=1
done
run_worker cube VH_MGS_MODE=64
run_worker mg VH_HALO_PUSH=0
run_worker cube VH_MGS_MODE=1000 VH_P2P=0
for pc in bj mg; do
  port=$((port + 1))
  tag="bench_n${N}_${pc}"
  ( timeout 600 $TR --master-port $port bench.py --gpus $N --steps 5 --warmup 1 --precond $pc "$@" 2> gpurun_out/$tag.err | tail -1 ) > gpurun_out/$tag.json
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/$tag.json"))
    print("$tag", "ms/step %.2f" % j["ms_per_step"], "its/step %.1f" % j["gmres_its_per_step"], "ms/it %.3f" % j["ms_per_gmres_it"],
          j["phase_ms_per_step"], "apply %.3f" % j["kernels"]["apply_ms"], "halo %.4f allreduce %.4f" % (j.get("halo_ms_per_exchange", -1), j.get("allreduce_ms_per_dot", -1)),
          "parity", j.get("multi_gpu_parity"))
except Exception as e:
    print("$tag ERR", e)
PY
done
# BASELINE configs[2]: Q2 r5 split over the GPUs
port=$((port + 1))
( timeout 600 $TR --master-port $port bench.py --gpus $N --steps 3 --warmup 1 --workload c3 2> gpurun_out/bench_n${N}_c3.err | tail -1 ) > gpurun_out/bench_n${N}_c3.json
python - <<PY
import json
try:
    j = json.load(open("gpurun_out/bench_n${N}_c3.json"))
    print("c3 n$N", "ms/step %.2f" % j["ms_per_step"], "its/step %.1f" % j["gmres_its_per_step"], "ms/it %.3f" % j["ms_per_gmres_it"], j["phase_ms_per_step"],
          "apply %.3f" % j["kernels"]["apply_ms"], "parity", j.get("multi_gpu_parity"))
except Exception as e:
    print("c3 ERR", e)
PY
