#!/bin/bash
# GPU visit r02g: parity suite; A/B old (HEAD~2: no TMA, per-pair atomics) / base (TMA tile clears + warp-aggregated atomics) /
# tmaepi (+ TMA block stores in the epilogue) / noagg (per-pair atomics) on c3 + c2 + c4; ncu of the binning kernels (atomics counters)
# for base / noagg on c2 and c4; ncu --set full of the c4 tile kernel
o=gpurun_out; mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -x -q > $o/pytest_gpu_r02g.log 2>&1; echo "pytest rc=$?"; tail -6 $o/pytest_gpu_r02g.log
REPS="1 2" EXTRA_CFG=c2,c4 tools/gpu_ab2.sh r02g c3 old base tmaepi noagg
M=lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed_op_global_atom.sum,smsp__inst_executed_op_global_red.sum,dram__bytes_read.sum,dram__bytes_write.sum
for v in base noagg; do for cfg in c2 c4; do
  lib=$PWD/ifrit-v2_b200/lib_ab/$v/libifrit_b200.so; [ $v = base ] && lib=$PWD/ifrit-v2_b200/lib/libifrit_b200.so
  IFRIT_B200_LIB=$lib timeout 600 ncu --metrics $M --clock-control none -k regex:'setup_count_kernel|fill_light_kernel|fill_heavy_kernel' -s 9 -c 3 --csv --log-file $o/atom_${v}_${cfg}_r02g.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-exact-sampler --extra-configs= --config $cfg > $o/atom_${v}_${cfg}_r02g.log 2>&1; echo "ncu atomics $v $cfg rc=$?"
done; done
tools/gpu_ncu.sh r02g c4 tile_kernel
