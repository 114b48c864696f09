#!/bin/bash
# usage (GPU box): tools/gpu_r2.sh <tag> [skip-tests] — round-2 visit: parity tests (incl. full-size configs), the default bench line
# (c3 headline + c2/c4/c5 extras + cpu baseline), launch list, ncu --set full of the tile kernel (c3, c2, c4) and of the
# geometry kernels on the soup (c4)
tag=${1:-x}; o=gpurun_out; mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $o/smi_$tag.txt 2>&1
nproc > $o/nproc_$tag.txt; free -g >> $o/nproc_$tag.txt
if [ "$2" != "skip-tests" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $o/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -14 $o/pytest_gpu_$tag.log
fi
timeout 1200 python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err; echo "bench rc=$?"; tail -2 $o/bench_$tag.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_c3_$tag.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fast-sampler --extra-configs= > $o/launches_c3_$tag.log 2>&1; echo "launch list rc=$?"
for cfg in c3 c2 c4; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 4 -c 1 -o $o/prof_tile_${cfg}_$tag -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fast-sampler --extra-configs= --config $cfg > $o/ncu_${cfg}_$tag.log 2>&1; echo "ncu tile $cfg rc=$?"
done
for k in setup_count_kernel fill_light_kernel fill_heavy_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o $o/prof_${k}_c4_$tag -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fast-sampler --extra-configs= --config c4 > $o/ncu_${k}_c4_$tag.log 2>&1; echo "ncu $k c4 rc=$?"
done
python - <<PY
import json
try:
    o=json.loads(open('$o/bench_$tag.json').read().strip().splitlines()[-1])
    def show(n,d): print(n, 'ms/step %.4f'%d['ms_per_step'], 'Mtris/s %.1f'%d['value'], 'kernel_ms', {k:round(v,4) for k,v in d['kernel_ms'].items()}, 'frac %.4f'%d['roofline']['frac'], ('e2e ms %.3f'%d['e2e']['ms_per_step']) if 'e2e' in d else '')
    show(o['config']['name'], o)
    for k,v in o.get('extra_configs',{}).items():
        if 'error' in v: print(k, v)
        else: show(k, v)
    print("exact", o.get("exact_sampler"))
    if 'cpu_baseline' in o: print('cpu', o['cpu_baseline'])
except Exception as e: print('parse fail', e)
PY
