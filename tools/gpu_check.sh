#!/bin/bash
# usage (on the GPU box): tools/gpu_check.sh [bench args...] — parity tests then a short bench of c2 + c3
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    o=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    def show(n,d): print(n, 'ms/step %.4f'%d['ms_per_step'], 'Mtris/s %.1f'%d['value'], 'kernel_ms', {k:round(v,4) for k,v in d['kernel_ms'].items()}, 'frac %.4f'%d['roofline']['frac'], 'e2e ms %.3f'%d['e2e']['ms_per_step'])
    show(o['config']['name'], o)
    for k,v in o.get('extra_configs',{}).items():
        if 'error' in v: print(k, v)
        else: show(k, v)
except Exception as e: print('parse fail', e)
PY
