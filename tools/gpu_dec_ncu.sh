#!/bin/bash
# ncu launch list of one decoding step (bf16 weights; the second of two eager steps) + full-set capture of its first kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "dec_step_199/" --csv --log-file gpurun_out/dec_launches.csv python tools/dec_step.py > gpurun_out/dec_ncu.log 2>&1
tail -2 gpurun_out/dec_ncu.log
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "dec_step_199/" -c 12 -o gpurun_out/dec_full -f python tools/dec_step.py >> gpurun_out/dec_ncu.log 2>&1
ncu -i gpurun_out/dec_full.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_tensor.sum,launch__grid_size,launch__registers_per_thread,launch__occupancy_limit_shared_mem > gpurun_out/dec_full_raw.csv 2>&1
ls -la gpurun_out | tail -5
