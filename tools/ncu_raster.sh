#!/bin/bash
# ncu --set full of the rasteriser kernels at the benchmark batch (64 views; dense, coarse and mixed meshes); the report
# is converted to a CSV summary on the box (gpurun_out/ travels back, the .ncu-rep does not).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
timeout 600 $NCU --set full --import-source on --clock-control none --kernel-name 'regex:k_raster' --launch-skip 6 \
  --launch-count 12 -f -o /tmp/r02_raster python tools/render_bench.py > gpurun_out/raster_prof.log 2>&1
tail -2 gpurun_out/raster_prof.log
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,launch__grid_size,launch__block_size'
$NCU -i /tmp/r02_raster.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_ncu_raster_summary.csv 2> gpurun_out/raster_prof3.log
echo "summary lines: $(wc -l < gpurun_out/r02_ncu_raster_summary.csv)"
