#!/bin/bash
# usage (GPU box): tools/gpu_round.sh <tag> — parity tests, full bench line, c4 bench, launch list, ncu --set full of the tile kernel (c2, c3)
tag=${1:-x}; o=gpurun_out; mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $o/smi_$tag.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $o/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $o/pytest_gpu_$tag.log
timeout 900 python bench.py > $o/bench_c2_$tag.json 2> $o/bench_c2_$tag.err; echo "bench c2 rc=$?"; tail -2 $o/bench_c2_$tag.err
timeout 600 python bench.py --config c4 --no-cpu-baseline --extra-configs= --steps 10 --warmup 3 > $o/bench_c4_$tag.json 2> $o/bench_c4_$tag.err; echo "bench c4 rc=$?"; tail -2 $o/bench_c4_$tag.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_c2_$tag.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extra-configs= > $o/launches_c2_$tag.log 2>&1; echo "launch list rc=$?"
for cfg in c2 c3; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 4 -c 1 -o $o/prof_tile_${cfg}_$tag -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --extra-configs= --config $cfg > $o/ncu_${cfg}_$tag.log 2>&1; echo "ncu $cfg rc=$?"
done
python - <<PY
import json
for c in ('c2','c4'):
    try:
        o=json.loads(open('$o/bench_%s_$tag.json'%c).read().strip().splitlines()[-1])
        def show(n,d): print(n, 'ms/step %.4f'%d['ms_per_step'], 'Mtris/s %.1f'%d['value'], 'kernel_ms', {k:round(v,4) for k,v in d['kernel_ms'].items()}, 'frac %.4f'%d['roofline']['frac'], 'e2e ms %.3f'%d['e2e']['ms_per_step'])
        show(o['config']['name'], o)
        for k,v in o.get('extra_configs',{}).items():
            if 'error' in v: print(k, v)
            else: show(k, v)
        if 'cpu_baseline' in o: print('cpu', o['cpu_baseline'])
    except Exception as e: print('parse fail', c, e)
PY
