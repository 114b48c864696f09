#!/bin/bash
# usage (N-GPU box): tools/gpu_dsetup2.sh [N] — bit identity incl. distributed setup + c5 with / without it
n=${1:-2}; o=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | grep -E "dsetup|MGPU|Error|error" | head -12
for f in "--force-distributed-setup" "--no-distributed-setup"; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n --config c5 --steps 5 --warmup 3 --no-cpu-baseline --no-exact-sampler --extra-configs= $f 2>$o/c5_n${n}_r03.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        o=json.loads(l); print('c5 n=$n $f', round(o['ms_per_step'],3), {k:round(v,3) for k,v in o['kernel_ms'].items()}, o.get('frame_hash_equals_1gpu'))
"; done; tail -2 $o/c5_n${n}_r03.err
