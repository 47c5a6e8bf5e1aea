#!/bin/bash
mkdir -p gpurun_out
echo "=== ncu launch list c3 (utts 24, no graph, warm caches)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1100 -c 1400 --csv --log-file gpurun_out/launches_c3.csv python bench.py --dtype bf16 --workload c3 --utts 24 --steps 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_list_c3.log 2>&1; tail -1 gpurun_out/ncu_list_c3.log | cut -c1-200
echo "=== graph-mode bench c3 utts 24, PDL on/off"
python bench.py --dtype bf16 --workload c3 --utts 24 --steps 10 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
CST_PDL=0 python bench.py --dtype bf16 --workload c3 --utts 24 --steps 10 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
python bench.py --dtype bf16 --workload c3 --utts 24 --steps 10 --no-cpu-baseline --no-graph 2>&1 | tail -1 | cut -c1-200
