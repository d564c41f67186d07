#!/bin/bash
# compute-sanitizer over the hot path: memcheck + racecheck + synccheck on smoke() (assembly, matrix-free
# GMRES with the cooperative Gram-Schmidt kernel, line search), and memcheck on a 2-rank run (peer-memory mailboxes, fused
# ghost push, NCCL halo).  Run under gpurun (--gpus 2 for the multi-rank part); logs land in gpurun_out/sanitizer_*.log.
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 900 $CS --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_smoke_$tool.log 2>&1
  echo "smoke $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok' gpurun_out/sanitizer_smoke_$tool.log | tr '\n' ' ')"
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  for push in 0 1; do
    VH_HALO_PUSH=$push timeout 1200 $CS --tool memcheck --target-processes all --print-limit 20 \
      python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700 + push)) tests/multigpu_worker.py \
      > gpurun_out/sanitizer_2rank_memcheck_push$push.log 2>&1
    echo "2-rank memcheck push=$push: $(grep -E 'ERROR SUMMARY|PARITY' gpurun_out/sanitizer_2rank_memcheck_push$push.log | sort | uniq -c | tr '\n' ' ')"
  done
fi
