#!/bin/bash
# usage (N-GPU box): tools/gpu_r04c.sh [N] — bit identity incl. the frame ring, then c3 at N with ring 2 / ring 1
n=${1:-2}; o=gpurun_out; mkdir -p $o
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | grep -E "mode=|MGPU|Error|error" | head -16 | tee $o/mgpu_check_r04c_n$n.log
for ring in ${RINGS:-2 1}; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n --steps 1000 --warmup 5 --no-cpu-baseline --no-exact-sampler --extra-configs=$EXTRA --ring $ring > $o/bench_c3_n${n}_ring${ring}_r04c.json 2> $o/bench_c3_n${n}_ring${ring}_r04c.err
tail -2 $o/bench_c3_n${n}_ring${ring}_r04c.err | cut -c1-300
python - <<PY
import json
for l in open('$o/bench_c3_n${n}_ring${ring}_r04c.json'):
    if l.startswith('{'):
        o=json.loads(l); print('c3 n=$n ring=$ring', round(o['ms_per_step'],4), {k:round(v,4) for k,v in o['kernel_ms'].items()}, 'ranks', o.get('tile_ms_per_rank'), 'hash', o.get('frame_hash_equals_1gpu'), 'e2e', round(o['e2e']['ms_per_step'],3), o['e2e'].get('frame_equals_1gpu'), 'gather', round((o.get('e2e_gather_rank0') or {}).get('ms_per_step',0),3))
        for k,x in (o.get('extra_configs') or {}).items(): print('  ', k, round(x.get('ms_per_step',0),3), x.get('kernel_ms'), x.get('frame_hash_equals_1gpu'))
PY
done
