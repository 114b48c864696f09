#!/bin/bash
# usage (1-GPU box): r04 evidence run — GPU tests, the default bench line, launch list, ncu --set full of the tile kernel (c3) and of the
# compacted setup pass (c5 as rank 0 of 8), MINB A/B of that pass, zero-copy vs DMA read-back
o=gpurun_out; mkdir -p $o; tag=r04
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $o/smi_$tag.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $o/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 $o/pytest_gpu_$tag.log
t0=$(date +%s); timeout 900 python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err; echo "bench rc=$? $(( $(date +%s) - t0 )) s"; tail -2 $o/bench_$tag.err | cut -c1-300
t0=$(date +%s); timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $o/bench_ref_$tag.json 2> $o/bench_ref_$tag.err; echo "bench ref rc=$? $(( $(date +%s) - t0 )) s"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_c3_$tag.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-exact-sampler --extra-configs= > $o/launches_c3_$tag.log 2>&1; echo "launch list rc=$?"
tools/gpu_ncu.sh $tag c3 tile_kernel
tools/gpu_ncu.sh $tag c5 setup_count_masked --emulate-rank 0/8 --force-distributed-setup
IFRIT_B200_LIB=$PWD/ifrit-v2_b200/lib_ab/minb4/libifrit_b200.so EMUL_ARGS=--force-distributed-setup tools/gpu_emul.sh c5 8
python - <<'PY' 2>&1 | tee gpurun_out/readback_r04.txt
# zero-copy stores (iftrcuCopyOwnedTilesToHost, one rank = whole frame) against the DMA copy (iftrcuCopyColorToHostAsync), 4K RGBA32F
import time, numpy as np, torch
import ifrit_v2_b200
from ifrit_v2_b200 import renderer as R, multigpu as MG
W, H = 3840, 2160
r = R.TileRasterRendererCuda(0); r.Init()
r.bindFrameBuffer(None, None, W, H); r.setClearValues([(0.1, 0.2, 0.3, 1)], 1.0); r.clear(); r.synchronize()
img = MG.SharedHostImage(W, H)
pin = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
for name, fn in (("zero-copy kernel -> shared mapped image", lambda: r.copyOwnedTilesToHost(img.ptr)), ("DMA -> pinned image", lambda: r.copyColorToHostAsync(pin.data_ptr()))):
    fn(); r.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): fn()
    r.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print(f"{name}: {dt * 1e3:.3f} ms per 133 MB frame = {W * H * 16 / dt / 1e9:.1f} GB/s")
print("equal:", bool(np.array_equal(img.array, pin.numpy())))
PY
python - <<PY
import json
o=json.loads(open('$o/bench_$tag.json').read().strip().splitlines()[-1])
def show(n,d): print(n, 'ms/step %.4f'%d['ms_per_step'], 'Mtris/s %.1f'%d['value'], 'kernel_ms', {k:round(v,4) for k,v in d['kernel_ms'].items()}, 'frac %.4f'%d['roofline']['frac'], 'alone', d['roofline'].get('frac_unpipelined'), 'e2e ms %.3f'%d['e2e']['ms_per_step'])
show(o['config']['name'], o)
for k,v in o.get('extra_configs',{}).items():
    if 'error' in v: print(k, v)
    else: show(k, v)
print('cpu', o.get('cpu_baseline')); print('exact', {k:o['exact_sampler'].get(k) for k in ('ms_per_step','value')} if 'exact_sampler' in o else None)
PY
