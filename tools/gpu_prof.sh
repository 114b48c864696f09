#!/bin/bash
# usage (GPU box): tools/gpu_prof.sh <config> <tag> [kernel regex] — one ncu --set full capture of the dominant kernel
cfg=${1:-c2}; tag=${2:-x}; k=${3:-tile_kernel}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/prof_${k}_${cfg}_${tag} -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --extra-configs= --config $cfg > gpurun_out/ncu_${cfg}_${tag}.log 2>&1
echo "ncu rc=$?"
