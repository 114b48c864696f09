#!/bin/bash
# usage (GPU box with N GPUs): tools/gpu_mgpu.sh <N> <tag> — sort-first bit-identity check + bench c2 / c5 at N GPUs
N=${1:-2}; tag=${2:-x}; o=gpurun_out; mkdir -p $o
nvidia-smi -L > $o/smi_mgpu_$tag.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py > $o/mgpu_check_$tag.log 2>&1; echo "mgpu_check rc=$?"; grep -E "identical|MGPU" $o/mgpu_check_$tag.log
for cfg in ${CFGS:-c3 c5}; do
for n in 1 $N; do
  if [ $n = 1 ]; then cmd="python"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544"; fi
  steps=200; [ $cfg = c5 ] && steps=5
  timeout 900 $cmd bench.py --gpus $n --config $cfg --steps $steps --warmup 3 --no-cpu-baseline --extra-configs= > $o/bench_${cfg}_n${n}_$tag.json 2> $o/bench_${cfg}_n${n}_$tag.err; echo "bench $cfg n=$n rc=$?"
  python - <<PY
import json
try:
    o=json.loads(open('$o/bench_${cfg}_n${n}_$tag.json').read().strip().splitlines()[-1])
    print('$cfg n=$n', 'ms/step %.4f'%o['ms_per_step'], 'Mtris/s %.1f'%o['value'], {k:round(v,4) for k,v in o['kernel_ms'].items()}, 'e2e ms %.3f'%o['e2e']['ms_per_step'])
except Exception as e: print('parse fail', e); print(open('$o/bench_${cfg}_n${n}_$tag.err').read()[-1500:])
PY
done; done
