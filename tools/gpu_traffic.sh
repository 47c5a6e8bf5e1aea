#!/bin/bash
# DRAM traffic per launch of every kernel of one c3 step (24 utterances = 4 batches), for roofline.traffic
mkdir -p gpurun_out
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 1400 -c 1400 --csv --log-file gpurun_out/traffic_c3.csv python bench.py --dtype bf16 --workload c3 --utts 24 --steps 1 --lanes 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1
tail -1 gpurun_out/ncu_traffic.log | cut -c1-100; wc -l gpurun_out/traffic_c3.csv
