#!/bin/bash
# GPU session (1 GPU): which C4 geometry / cycle threshold gives a Newton run that converges at the final mesh
set -u
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 400 python tools/c4_adaptive.py "$@" --json gpurun_out/s14_c4_$tag.json > gpurun_out/s14_c4_$tag.log 2>&1; echo "== $tag: $*"; grep -E "rank\(s\)\]|^cycle|Error" gpurun_out/s14_c4_$tag.log | cut -c1-170 | tail -22; }
run a --half 6 4 8 --initial-refine 4 --threshold 1.0
run b --half 3 2 4 --initial-refine 4 --threshold 1e3
run c --half 3 2 4 --initial-refine 4 --threshold 1.0
