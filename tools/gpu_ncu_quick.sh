#!/bin/bash
# targeted ncu captures (c2 batch, eager): conv0_tc, one w2v layer's GEMMs
mkdir -p gpurun_out
B="python bench.py --dtype bf16 --workload c2 --steps 1 --lanes 1 --no-graph --no-cpu-baseline"
echo "=== conv0_tc"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv0_tc -s 1 -c 1 -f -o gpurun_out/prof_conv0 $B > gpurun_out/ncu_conv0.log 2>&1; tail -1 gpurun_out/ncu_conv0.log | cut -c1-150
echo "=== gemm"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 0 -c 15 -f -o gpurun_out/prof_gemm $B > gpurun_out/ncu_gemm.log 2>&1; tail -1 gpurun_out/ncu_gemm.log | cut -c1-150
ls -la gpurun_out/*.ncu-rep
