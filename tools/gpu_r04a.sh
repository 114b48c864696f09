#!/bin/bash
# usage (1-GPU box): r04a — GPU tests, then the masked (compacted) setup pass on c5 as rank 0 of 8 / of 4 (1-GPU emulation), two builds
o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $o/pytest_gpu_r04a.log
for v in base minb3; do
  lib=$PWD/ifrit-v2_b200/lib/libifrit_b200.so; [ $v = minb3 ] && lib=$PWD/ifrit-v2_b200/lib_ab/minb3/libifrit_b200.so
  echo "== $v"; IFRIT_B200_LIB=$lib EMUL_ARGS=--force-distributed-setup tools/gpu_emul.sh c5 8
done
EMUL_ARGS=--force-distributed-setup tools/gpu_emul.sh c5 4
tools/gpu_emul.sh c5 1
