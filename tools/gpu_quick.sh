#!/bin/bash
# usage (GPU box): tools/gpu_quick.sh <tag> [extra configs] [pytest -k expr] — parity tests, then a short bench (c3 headline + extras), no CPU leg
tag=${1:-x}; extra=${2:-c2,c4}; o=gpurun_out; mkdir -p $o
if [ -n "$3" ]; then timeout 1500 python -m pytest tests -m gpu -x -q -k "$3" > $o/pytest_gpu_$tag.log 2>&1; else timeout 1500 python -m pytest tests -m gpu -x -q > $o/pytest_gpu_$tag.log 2>&1; fi
echo "pytest rc=$?"; tail -6 $o/pytest_gpu_$tag.log
timeout 900 python bench.py --steps 300 --no-cpu-baseline --extra-configs=$extra > $o/bench_$tag.json 2> $o/bench_$tag.err; echo "bench rc=$?"; tail -3 $o/bench_$tag.err
python - <<PY
import json
try:
    o=json.loads(open('$o/bench_$tag.json').read().strip().splitlines()[-1])
    def show(n,d): print(n, 'ms/step %.4f'%d['ms_per_step'], 'kernel_ms', {k:round(v,4) for k,v in d['kernel_ms'].items()}, 'alone', {k:round(v,4) for k,v in (d.get('kernel_ms_unpipelined') or {}).items()}, 'frac %.4f'%d['roofline']['frac'], ('e2e ms %.3f'%d['e2e']['ms_per_step']) if 'e2e' in d else '')
    show(o['config']['name'], o)
    for k,v in o.get('extra_configs',{}).items():
        if 'error' in v: print(k, v)
        else: show(k, v)
    x=o.get('exact_sampler')
    if x: print('exact', x.get('ms_per_step'), x.get('kernel_ms'))
except Exception as e: print('parse fail', e)
PY
