#!/bin/bash
# ncu evidence of round 2, sized to travel back from the GPU box (gpurun_out/ <= 64 MiB): the reports are converted to
# CSV on the box and deleted.  1) DRAM bytes + duration of every kernel of one forward of 64 hypotheses (+ one multiview
# scene); 2) `--set full` of the kernels the design rests on, a few launches each.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
timeout 300 $NCU --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  --clock-control none --csv --log-file gpurun_out/r02_ncu_dram_per_forward.csv python tools/prof_forward.py --multiview \
  > gpurun_out/prof1.log 2>&1
echo "dram csv lines: $(wc -l < gpurun_out/r02_ncu_dram_per_forward.csv)"
[ "$1" = "--dram-only" ] && exit 0
timeout 900 $NCU --profile-from-start off --set full --import-source on --clock-control none \
  --kernel-name 'regex:k_xdw|k_pw2|k_stem|k_roi_crop|k_se_gate|k_se_fc2|k_dw_tile|k_dwconv_roll|k_pool_fc|k_ransac|k_ba_|k_lm_solve|k_vote' \
  --launch-count 70 -f -o /tmp/r02_full python tools/prof_forward.py --multiview > gpurun_out/prof2.log 2>&1
tail -2 gpurun_out/prof2.log
ls -la /tmp/r02_full.ncu-rep
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size'
$NCU -i /tmp/r02_full.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_ncu_full_summary.csv 2> gpurun_out/prof3.log
echo "summary lines: $(wc -l < gpurun_out/r02_ncu_full_summary.csv)"
$NCU -i /tmp/r02_full.ncu-rep --page details --csv > /tmp/details.csv 2>/dev/null
grep -E "k_xdw|k_pw2" /tmp/details.csv | grep -E "Stall|Issue|Pipe|Throughput|Theoretical|Achieved" | head -400 > gpurun_out/r02_ncu_details_xdw_pw2.csv
ls -la gpurun_out/ | tail -8
