#!/bin/bash
# usage (GPU box): tools/gpu_ncu.sh <tag> <cfg> [kernel regex] [extra bench args] — one ncu --set full capture (+ source counters) of one kernel
tag=${1:-x}; cfg=${2:-c3}; k=${3:-tile_kernel}; shift $(( $# < 3 ? $# : 3 )); o=gpurun_out; mkdir -p $o
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o $o/prof_${k}_${cfg}_$tag -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-exact-sampler --extra-configs= --config $cfg "$@" > $o/ncu_${k}_${cfg}_$tag.log 2>&1; echo "ncu $k $cfg rc=$?"
