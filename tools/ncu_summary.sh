#!/bin/bash
# usage (here): tools/ncu_summary.sh <report.ncu-rep> <object file holding the kernel> <mangled kernel> <out.txt> [topN]
# headline counters (--page raw) + per-source-line stall / instruction table (tools/ncu_lines.py) of one ncu --set full capture
rep=$1; obj=$2; kern=$3; out=$4; top=${5:-30}
d=$(mktemp -d); (cd $d && cuobjdump -xelf all $obj > /dev/null 2>&1); cub=$(for c in $d/*.cubin; do if cuobjdump -elf $c 2>/dev/null | grep -q "$kern"; then echo $c; break; fi; done)
{
echo "# $(basename $rep): ncu --set full --clock-control none --import-source on, one launch; kernel $kern"
ncu -i $rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
keep=('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__shared_mem_per_block_static','launch__shared_mem_per_block_dynamic','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active','smsp__thread_inst_executed_per_inst_executed.ratio','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','smsp__inst_executed_op_texture.sum','smsp__inst_executed_op_shared_atom.sum','smsp__inst_executed_op_global_atom.sum','smsp__inst_executed_op_global_red.sum','lts__t_sectors_op_atom.sum','lts__t_sectors_op_red.sum','smsp__inst_executed_op_tma_st.sum','smsp__inst_executed_op_tma_ld.sum','Kernel Name')
for n,uu,vv in zip(h,u,v):
    if n in keep: print(f'{n},{uu},{vv}')
"
python3 /root/repo/tools/ncu_lines.py $rep $cub $kern $top 2>&1
} > $out
rm -rf $d; head -5 $out
