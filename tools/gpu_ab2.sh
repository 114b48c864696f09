#!/bin/bash
# usage (GPU box): tools/gpu_ab2.sh <tag> <cfg> <variant>... — kernel_ms of build variants (ifrit-v2_b200/lib_ab/<v>/libifrit_b200.so; "base" =
# ifrit-v2_b200/lib) back to back on one box; REPS (default "1 2") repetitions; EXTRA_CFG = further configs per run
tag=$1; cfg=$2; shift 2; o=gpurun_out; mkdir -p $o
for rep in ${REPS:-1 2}; do for v in "$@"; do
  lib=$PWD/ifrit-v2_b200/lib_ab/$v/libifrit_b200.so; [ $v = base ] && lib=$PWD/ifrit-v2_b200/lib/libifrit_b200.so
  IFRIT_B200_LIB=$lib timeout 600 python bench.py --config $cfg --steps 300 --warmup 3 --no-cpu-baseline --no-exact-sampler --extra-configs=$EXTRA_CFG > $o/ab_${tag}_${v}_$rep.json 2> $o/ab_${tag}_${v}_$rep.err
  python - <<PY
import json
try:
    o=json.loads(open('$o/ab_${tag}_${v}_$rep.json').read().strip().splitlines()[-1])
    def show(n,d): print('$v rep$rep', n, 'ms/step %.4f'%d['ms_per_step'], 'tile %.4f'%d['kernel_ms']['tile'], 'alone', {k:round(x,4) for k,x in (d.get('kernel_ms_unpipelined') or {}).items()})
    show(o['config']['name'], o)
    for k,x in o.get('extra_configs',{}).items():
        if 'error' in x: print(k, x)
        else: show(k, x)
except Exception as e: print('$v parse fail', e); print(open('$o/ab_${tag}_${v}_$rep.err').read()[-800:])
PY
done; done
