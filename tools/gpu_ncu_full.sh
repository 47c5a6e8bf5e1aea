#!/bin/bash
# ncu captures on the c2 batch (B=32 x 15 s, bf16), eager launches: launch list + --set full of the top kernels
mkdir -p gpurun_out
B="python bench.py --dtype bf16 --workload c2 --steps 1 --lanes 1 --no-graph --no-cpu-baseline"
echo "=== launch list (one step = 170 launches; the first 3 warm-up steps are skipped)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 510 -c 170 --csv --log-file gpurun_out/launches_c2.csv $B > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log | cut -c1-120
echo "=== gemm (conv1..6, proj, then layer 0/1 qkv,out,fc1,fc2)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 0 -c 15 -f -o gpurun_out/prof_gemm $B > gpurun_out/ncu_gemm.log 2>&1; tail -1 gpurun_out/ncu_gemm.log
echo "=== attn"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -f -o gpurun_out/prof_attn $B > gpurun_out/ncu_attn.log 2>&1; tail -1 gpurun_out/ncu_attn.log
echo "=== conv0"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv0_tc -s 1 -c 1 -f -o gpurun_out/prof_conv0 $B > gpurun_out/ncu_conv0.log 2>&1; tail -1 gpurun_out/ncu_conv0.log
echo "=== posconv"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:posconv_tc -s 1 -c 1 -f -o gpurun_out/prof_posconv $B > gpurun_out/ncu_posconv.log 2>&1; tail -1 gpurun_out/ncu_posconv.log
echo "=== ln"; timeout 900 ncu --set full --clock-control none -k regex:layernorm_kernel -s 3 -c 1 -f -o gpurun_out/prof_ln $B > gpurun_out/ncu_ln.log 2>&1; tail -1 gpurun_out/ncu_ln.log
ls -la gpurun_out/*.ncu-rep
