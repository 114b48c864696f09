#!/bin/bash
# usage (2-GPU box): r04b — bit identity incl. distributed (compacted) setup, then c3 at N = 2 with the distributed read-back in e2e
o=gpurun_out; mkdir -p $o
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | grep -E "mode=|MGPU|Error|error" | head -14 | tee $o/mgpu_check_r04b.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu-baseline --no-exact-sampler --extra-configs= > $o/bench_c3_n2_r04b.json 2> $o/bench_c3_n2_r04b.err
tail -3 $o/bench_c3_n2_r04b.err
python - <<PY
import json
for l in open('$o/bench_c3_n2_r04b.json'):
    if l.startswith('{'):
        o=json.loads(l); print('c3 n=2', round(o['ms_per_step'],4), o['kernel_ms'], 'hash', o.get('frame_hash_equals_1gpu')); print('e2e', o.get('e2e')); print('gather', o.get('e2e_gather_rank0'))
PY
