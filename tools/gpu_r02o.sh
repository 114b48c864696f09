#!/bin/bash
# GPU visit r02o: parity suite, then A/B of the tile-kernel grid hint (exact CTA count vs one CTA per tile) on c3 / c2 / c4
o=gpurun_out; mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -q > $o/pytest_gpu_r02o.log 2>&1; echo "pytest rc=$?"; tail -5 $o/pytest_gpu_r02o.log
for rep in 1 2; do for v in hint nohint; do
  flag=""; [ $v = nohint ] && flag="--no-grid-hint"
  python bench.py --config c3 --steps 300 --warmup 3 --no-cpu-baseline --no-exact-sampler --extra-configs=c2,c4 $flag > $o/ab_r02o_${v}_$rep.json 2> $o/ab_r02o_${v}_$rep.err
  python - <<PY
import json
o=json.loads(open('$o/ab_r02o_${v}_$rep.json').read().strip().splitlines()[-1])
x=o['extra_configs']
print('$v rep$rep c3 step %.4f tile %.4f alone %.4f | c2 step %.4f tile %.4f alone %.4f | c4 step %.4f tile alone %.4f' % (o['ms_per_step'], o['kernel_ms']['tile'], o['kernel_ms_unpipelined']['tile'], x['c2']['ms_per_step'], x['c2']['kernel_ms']['tile'], x['c2']['kernel_ms_unpipelined']['tile'], x['c4']['ms_per_step'], x['c4']['kernel_ms_unpipelined']['tile']))
PY
done; done
