#!/bin/bash
# GPU visit r02f: parity suite with the TMA framebuffer path, then A/B of the tile kernel: old (HEAD~, no TMA) / base (TMA) / tap4 / base --no-tma
o=gpurun_out; mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -x -q > $o/pytest_gpu_r02f.log 2>&1; echo "pytest rc=$?"; tail -6 $o/pytest_gpu_r02f.log
REPS="1 2" EXTRA_CFG=c2 tools/gpu_ab2.sh r02f c3 old base tap4
for rep in 1 2; do
python bench.py --config c3 --steps 300 --warmup 3 --no-cpu-baseline --no-exact-sampler --no-tma --extra-configs=c2 > $o/ab_r02f_notma_$rep.json 2> $o/ab_r02f_notma_$rep.err
python - <<PY
import json
o=json.loads(open('$o/ab_r02f_notma_$rep.json').read().strip().splitlines()[-1])
print('notma rep$rep c3 ms/step %.4f tile %.4f alone %.4f | c2 tile %.4f alone %.4f' % (o['ms_per_step'], o['kernel_ms']['tile'], o['kernel_ms_unpipelined']['tile'], o['extra_configs']['c2']['kernel_ms']['tile'], o['extra_configs']['c2']['kernel_ms_unpipelined']['tile']))
PY
done
