#!/bin/bash
# usage (here): tools/prof_report.sh <report.ncu-rep> <mangled kernel> [topN] — headline counters + per-line hotspots
rep=$1; kern=$2; top=${3:-40}
mkdir -p /tmp/sass && (cd /tmp/sass && rm -f *.cubin && cuobjdump -xelf all /root/repo/ifrit-v2_b200/lib/libifrit_b200.so > /dev/null)
cub=$(grep -l "$kern" /tmp/sass/*.cubin | head -1)
ncu -i $rep --page raw --csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
keep=('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__warps_eligible.avg.per_cycle_active','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active')
for n,uu,vv in zip(h,u,v):
    if n in keep: print(f'{n},{uu},{vv}')
"
python /root/repo/tools/ncu_lines.py $rep $cub $kern $top
