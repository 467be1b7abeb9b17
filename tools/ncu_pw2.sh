#!/bin/bash
# ncu --set full of every k_pw2 launch of one forward of 64 hypotheses (44 launches), summary CSV only.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
timeout 600 $NCU --profile-from-start off --set full --import-source on --clock-control none --kernel-name 'regex:k_pw2' \
  --launch-count 44 -f -o /tmp/r02_pw2 python tools/prof_forward.py > gpurun_out/pw2_prof.log 2>&1
tail -2 gpurun_out/pw2_prof.log
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size'
$NCU -i /tmp/r02_pw2.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_ncu_pw2_splitk_summary.csv 2> gpurun_out/pw2_prof3.log
echo "summary lines: $(wc -l < gpurun_out/r02_ncu_pw2_splitk_summary.csv)"
