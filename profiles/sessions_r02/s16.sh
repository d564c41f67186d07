#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_zmultigrid.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -4
run() { tag=$1; shift; timeout 300 python tools/c4_adaptive.py "$@" --json gpurun_out/s16_c4_$tag.json > gpurun_out/s16_c4_$tag.log 2>&1; echo "== $tag: $*"; grep -E "rank\(s\)\]|^cycle|Error" gpurun_out/s16_c4_$tag.log | cut -c1-170 | tail -14; }
run ch8 --half 3 2 4 --initial-refine 4 --threshold 1.0 --cheb-degree 8
run ch4 --half 3 2 4 --initial-refine 4 --threshold 1.0 --cheb-degree 4 --cheb-range 10
run ch8r100 --half 3 2 4 --initial-refine 4 --threshold 1.0 --cheb-degree 8 --restart 100
