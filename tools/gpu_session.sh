#!/bin/bash
# One standard GPU session (run under gpurun from the repo root): tests, the two bench modes, the ncu launch list of the
# bench command and one full capture of the dominant kernels.  Everything lands in gpurun_out/; turn the raw files into
# the committed summaries with profiles/summarize.py (see its header).
#   gpurun --timeout 900 -- 'bash tools/gpu_session.sh [tests|bench|ncu|all]'
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu_session.sh multi 2'      (multi-rank checks; N = 2, 4 or 8; never under ncu)
# ncu replays every profiled launch ~40x: the captures below use a SHORT bench command (2 steps) and are limited with -c.
set -u
what=${1:-all}
mkdir -p gpurun_out
if [ "$what" = tests ] || [ "$what" = all ]; then
  timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
  tail -5 gpurun_out/gpu_tests.log
fi
if [ "$what" = bench ] || [ "$what" = all ]; then
  timeout 300 python bench.py > gpurun_out/bench_packed.json 2> gpurun_out/bench_packed.err
  timeout 300 python bench.py --spmv-mf --no-cpu-baseline > gpurun_out/bench_mf.json 2> gpurun_out/bench_mf.err
  timeout 120 python tools/time_spmv_modes.py 1 5 20 > gpurun_out/spmv_modes_q1_r5.txt 2>&1
  timeout 300 python tools/bench_q2.py 4 5 > gpurun_out/q2_timings.txt 2>&1
  python - <<'PY'
import json
for f in ("bench_packed", "bench_mf"):
    try:
        j = json.load(open("gpurun_out/%s.json" % f))
        print(f, "ms/step %.3f" % j["ms_per_step"], "e2e %.3f" % j["e2e"]["ms_per_step"], j["phase_ms_per_step"], j["kernels"])
    except Exception as e:
        print(f, "ERR", e)
PY
  cat gpurun_out/spmv_modes_q1_r5.txt gpurun_out/q2_timings.txt
fi
if [ "$what" = ncu ] || [ "$what" = all ]; then
  B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
  # every launch of the bench command with its device time (shares of the step)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_packed.csv $B \
    > gpurun_out/ncu_launches_packed.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_mf.csv $B --spmv-mf \
    > gpurun_out/ncu_launches_mf.log 2>&1
  # the dominant kernels, full metric set with source (3 launches each, after the warm-up launches)
  timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_spmv_sym18|k_rows_fast_q1|k_gather_apply' -s 6 -c 6 \
    -o gpurun_out/prof_packed -f $B > gpurun_out/ncu_full_packed.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_points' -s 8 -c 6 \
    -o gpurun_out/prof_mf -f $B --spmv-mf > gpurun_out/ncu_full_mf.log 2>&1
  ls -la gpurun_out/*.ncu-rep gpurun_out/launches_*.csv 2>/dev/null
fi
if [ "$what" = multi ]; then
  N=${2:-2}
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  port=29600
  for mode in cube periodic hanging; do   # parity of the N-rank run with the 1-rank run (tests/multigpu_worker.py)
    for env in "" "VH_SPMV_MF=1" "VH_HALO_PUSH=1" "VH_SPMV_MF=1 VH_HALO_PUSH=1"; do
      port=$((port + 1))
      arg=$mode; [ "$mode" = cube ] && arg=""
      ( env $env timeout 120 $TR --master-port $port tests/multigpu_worker.py $arg 2>&1 | grep -E "PARITY|Error|error" | tail -2 ) > gpurun_out/mg_${mode}_$(echo "$env" | tr ' =' '__').log
      echo "$mode [$env]: $(tail -1 gpurun_out/mg_${mode}_$(echo "$env" | tr ' =' '__').log)"
    done
  done
  for flags in "" "--spmv-mf"; do
    for env in "" "VH_HALO_PUSH=1"; do
      port=$((port + 1))
      tag=$(echo "n${N}${flags}_${env}" | tr ' =-' '___')
      ( env $env timeout 200 $TR --master-port $port bench.py --gpus $N --steps 5 --warmup 3 $flags 2> gpurun_out/bench_$tag.err | tail -1 ) > gpurun_out/bench_$tag.json
      python -c "import json,sys; j=json.load(open('gpurun_out/bench_$tag.json')); print('$tag', 'ms/step %.3f' % j['ms_per_step'], j['phase_ms_per_step'], [h['gmres_its'] for h in j['newton_history']])" || echo "$tag ERR"
    done
  done
fi
