#!/bin/bash
set -u
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python tools/c4_adaptive.py "$@" --json gpurun_out/s15_c4_$tag.json > gpurun_out/s15_c4_$tag.log 2>&1; echo "== $tag: $*"; grep -E "rank\(s\)\]|^cycle|Error" gpurun_out/s15_c4_$tag.log | cut -c1-170 | tail -16; }
run b100 --half 3 2 4 --initial-refine 4 --threshold 1e3 --restart 100 --max-lin-it 20000
run c100 --half 3 2 4 --initial-refine 4 --threshold 1.0 --restart 100 --max-lin-it 20000
run b3 --half 3 2 4 --initial-refine 3 --threshold 1e3
