#!/bin/bash
# usage (8-GPU box): tools/gpu_mgpu8.sh <tag> — bit identity at 8 ranks, c3 at N = 1/2/4/8 (pre-clear on; off at 8), c5 at N = 8, NVLink byte counters around the c3 N=8 run
tag=${1:-x}; o=gpurun_out; mkdir -p $o
run() { n=$1; shift; if [ $n = 1 ]; then python bench.py --gpus 1 "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n "$@"; fi; }
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    o=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'ms/step %.4f'%o['ms_per_step'], 'Mtris/s %.1f'%o['value'], {k:round(v,4) for k,v in o['kernel_ms'].items()}, 'hash', o.get('frame_hash_equals_1gpu'), 'e2e', round(o['e2e']['ms_per_step'],3) if 'e2e' in o else None)
except Exception as e: print(sys.argv[2], 'parse fail', e)
PY
}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py > $o/mgpu_check8_$tag.log 2>&1; echo "mgpu_check rc=$?"; grep -E "identical|MGPU" $o/mgpu_check8_$tag.log
for n in 1 2 4 8; do
  timeout 600 bash -c "$(declare -f run); run $n --config c3 --steps 300 --warmup 5 --no-cpu-baseline --no-exact-sampler --extra-configs=" > $o/bench_c3_n${n}_$tag.json 2> $o/bench_c3_n${n}_$tag.err; show $o/bench_c3_n${n}_$tag.json "c3 n=$n"
done
nvidia-smi nvlink -gt d -i 0 > $o/nvlink_before_$tag.txt 2>&1
timeout 600 bash -c "$(declare -f run); run 8 --config c3 --steps 2000 --warmup 5 --no-cpu-baseline --no-exact-sampler --extra-configs= --no-preclear" > $o/bench_c3_n8_nopre_$tag.json 2> $o/bench_c3_n8_nopre_$tag.err; show $o/bench_c3_n8_nopre_$tag.json "c3 n=8 no-preclear (2000 steps)"
nvidia-smi nvlink -gt d -i 0 > $o/nvlink_mid_$tag.txt 2>&1
timeout 600 bash -c "$(declare -f run); run 8 --config c3 --steps 2000 --warmup 5 --no-cpu-baseline --no-exact-sampler --extra-configs=" > $o/bench_c3_n8_pre_$tag.json 2> $o/bench_c3_n8_pre_$tag.err; show $o/bench_c3_n8_pre_$tag.json "c3 n=8 preclear (2000 steps)"
nvidia-smi nvlink -gt d -i 0 > $o/nvlink_after_$tag.txt 2>&1
for n in 1 8; do
  timeout 900 bash -c "$(declare -f run); run $n --config c5 --steps 5 --warmup 3 --no-cpu-baseline --no-exact-sampler --extra-configs=" > $o/bench_c5_n${n}_$tag.json 2> $o/bench_c5_n${n}_$tag.err; show $o/bench_c5_n${n}_$tag.json "c5 n=$n"
done
