#!/bin/bash
o=gpurun_out; mkdir -p $o
MALLOC_CHECK_=3 timeout 300 python tools/dbg_compat.py > $o/dbg_compat_r02h.log 2>&1; echo "dbg rc=$?"; tail -25 $o/dbg_compat_r02h.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_compat_abi.py::test_scene_built_through_handles_matches_pointer_path > $o/pytest_gpu_r02h.log 2>&1; echo "pytest rc=$?"; tail -8 $o/pytest_gpu_r02h.log
REPS="1 2" EXTRA_CFG=c2,c4 tools/gpu_ab2.sh r02h c3 old base hot
