#!/bin/bash
# usage (1-GPU box): tools/gpu_emul.sh <cfg> <N>... — per-rank kernel times (GPU to itself) of rank 0 of N for each N, emulated on one GPU
cfg=$1; shift; o=gpurun_out
for n in "$@"; do
timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-exact-sampler --extra-configs= --emulate-rank 0/$n $EMUL_ARGS 2>$o/emul_${cfg}_$n.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        o=json.loads(l); print('$cfg rank 0/$n step', round(o['ms_per_step'],3), 'alone', {k:round(v,3) for k,v in (o.get('kernel_ms_unpipelined') or {}).items()}, 'pipelined', {k:round(v,3) for k,v in o['kernel_ms'].items()})
"; done
