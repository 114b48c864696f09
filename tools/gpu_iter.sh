#!/bin/bash
# usage (GPU box): tools/gpu_iter.sh <tag> [configs...] — parity tests + short benches (no CPU baseline) for a perf iteration
tag=${1:-x}; shift; cfgs=${@:-c2 c4 c5}; o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -x -q > $o/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 $o/pytest_gpu_$tag.log
for cfg in $cfgs; do
  steps=300; extra=; [ $cfg = c2 ] && extra=c3; [ $cfg = c4 ] && steps=20; [ $cfg = c5 ] && steps=5
  timeout 900 python bench.py --config $cfg --steps $steps --warmup 3 --no-cpu-baseline --extra-configs=$extra > $o/bench_${cfg}_$tag.json 2> $o/bench_${cfg}_$tag.err; echo "bench $cfg rc=$?"
  python - <<PY
import json
try:
    o=json.loads(open('$o/bench_${cfg}_$tag.json').read().strip().splitlines()[-1])
    def show(n,d): print(n, "ms/step %.4f"%d["ms_per_step"], "lat %.4f"%(d.get("frame_latency_ms") or 0), "Mtris/s %.1f"%d["value"], {k:round(v,4) for k,v in (d.get("kernel_ms_unpipelined") or d["kernel_ms"]).items()}, "frac %.4f"%d["roofline"]["frac"], "e2e ms %.3f"%d["e2e"]["ms_per_step"])
    show(o['config']['name'], o)
    for k,v in o.get('extra_configs',{}).items():
        if 'error' in v: print(k, v)
        else: show(k, v)
except Exception as e: print('parse fail', e); print(open('$o/bench_${cfg}_$tag.err').read()[-1500:])
PY
done
