#!/bin/bash
# usage (2-GPU box): bit identity incl. distributed setup + c5 with / without it
o=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | grep -E "dsetup|MGPU|Error|error" | head -12
for f in "" "--no-distributed-setup"; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --config c5 --steps 5 --warmup 3 --no-cpu-baseline --no-exact-sampler --extra-configs= $f 2>$o/c5_n2_r02y.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        o=json.loads(l); print('c5 n=2 $f', round(o['ms_per_step'],3), {k:round(v,3) for k,v in o['kernel_ms'].items()}, o.get('frame_hash_equals_1gpu'))
"; done; tail -2 $o/c5_n2_r02y.err
