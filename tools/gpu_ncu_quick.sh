#!/bin/bash
# partial re-capture after a kernel change: launch list + one kernel (default: the pos-conv)
mkdir -p gpurun_out
K=${1:-posconv}
B="python bench.py --dtype bf16 --workload c2 --steps 1 --lanes 1 --no-graph --no-cpu-baseline"
echo "=== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 510 -c 170 --csv --log-file gpurun_out/launches_c2.csv $B > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log | cut -c1-100
echo "=== $K"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_posconv $B > gpurun_out/ncu_posconv.log 2>&1; tail -1 gpurun_out/ncu_posconv.log | cut -c1-100
