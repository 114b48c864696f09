#!/bin/bash
# usage (GPU box): tools/gpu_sanitize.sh <tag> — compute-sanitizer memcheck + racecheck over the GPU parity suite (full-size and NVRTC tests left out: minutes each under the tool)
tag=${1:-x}; o=gpurun_out; mkdir -p $o
K='not full_size and not runtime and not cpp_host and not reference_typed and not unsynchronised and not overflow'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 86 --print-limit 20 python -m pytest tests -m gpu -q -x -k "$K" > $o/sanitizer_memcheck_$tag.log 2>&1; echo "memcheck rc=$?"; tail -6 $o/sanitizer_memcheck_$tag.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 86 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "random_triangles or hierarchical or blend or late_z or point_and_line or tma or textured_mesh or beyond_4096" > $o/sanitizer_racecheck_$tag.log 2>&1; echo "racecheck rc=$?"; tail -6 $o/sanitizer_racecheck_$tag.log
