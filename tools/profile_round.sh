#!/bin/bash
# GPU-box profiling passes for one round (run under gpurun from the repo root):
#   1. launch list of the bench command (gpu__time_duration only, clocks untouched)
#   2. selected metrics for every launch of one training step
#   3. ncu --set full captures of the named kernels (source-level, -lineinfo)
# Outputs under gpurun_out/; summaries are made with tools/ncu_summary.py / ncu_raw_summary.py.
set -u
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --quick --steps 1 --warmup 3 --no-predict --no-cpu"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
  --log-file $OUT/launches.csv $BENCH > $OUT/launches.log 2>&1
M=gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
M=$M,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
M=$M,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,launch__grid_size,launch__registers_per_thread
M=$M,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
# one step = launches of the 4th pass; skip is computed from the launch list (launches per step)
N=$(python tools/ncu_summary.py $OUT/launches.csv --count 2>/dev/null || echo 0)
PER=$((N / 4))
if [ "$PER" -gt 0 ]; then
  timeout 900 ncu --metrics $M --clock-control none --launch-skip $((N - PER)) -c $PER \
    -o $OUT/step_metrics -f $BENCH > $OUT/step_metrics.log 2>&1
  ncu -i $OUT/step_metrics.ncu-rep --page raw --csv > $OUT/step_metrics_raw.csv 2>/dev/null
fi
for K in "$@"; do
  # K = kernel-regex:skip:count, e.g. conv_halo_kernel:40:2, conv_gemm2_kernel:150:4
  # (round 2: conv_gemm2_kernel:150:3 conv_halo_kernel:68:3 wgrad_gemm2_kernel:55:2 wgrad_halo_kernel:32:2 wgrad_gemm_kernel:26:2)
  IFS=: read -r NAME SKIP CNT <<< "$K"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$NAME \
    --launch-skip $SKIP -c $CNT -o $OUT/full_$NAME -f $BENCH > $OUT/full_$NAME.log 2>&1
  ncu -i $OUT/full_$NAME.ncu-rep --page raw --csv > $OUT/full_${NAME}_raw.csv 2>/dev/null
done
ls -la $OUT
