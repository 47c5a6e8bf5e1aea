#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest ops"; timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 600 2>&1 | tail -30 | tee gpurun_out/pytest_ops.log
echo "=== pytest encoder"; timeout 1200 python -m pytest tests/test_gpu_encoder.py -m gpu -q --timeout 900 2>&1 | tail -40 | tee gpurun_out/pytest_enc.log
echo "=== bf16 err"; timeout 600 python tools/bf16_err.py 2>&1 | tail -8 | tee gpurun_out/bf16_err.log
echo "=== bench bf16 c2"; timeout 600 python bench.py --dtype bf16 --workload c2 --steps 3 --no-cpu-baseline --profile-json gpurun_out/prof_c2.json 2>&1 | tail -2 | tee gpurun_out/bench_bf16_c2.log
echo "=== bench bf16 c3"; timeout 900 python bench.py --dtype bf16 --workload c3 --utts 256 --steps 2 --no-cpu-baseline --profile-json gpurun_out/prof_c3.json 2>&1 | tail -2 | tee gpurun_out/bench_bf16_c3.log
echo "=== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_c2.csv python bench.py --dtype bf16 --workload c2 --steps 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log
echo "=== ncu full gemm"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 30 -c 3 -f -o gpurun_out/prof_gemm python bench.py --dtype bf16 --workload c2 --steps 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
echo "=== ncu full conv0"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv0_apply -s 1 -c 1 -f -o gpurun_out/prof_conv0 python bench.py --dtype bf16 --workload c2 --steps 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_conv0.log 2>&1; tail -2 gpurun_out/ncu_conv0.log
echo "=== ncu full attn"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -f -o gpurun_out/prof_attn python bench.py --dtype bf16 --workload c2 --steps 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log
ls -la gpurun_out
