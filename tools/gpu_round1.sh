#!/bin/bash
# First GPU bring-up: diagnostics, parity tests (each file in its own process), short benches, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== tc_diag"; timeout 900 python tools/tc_diag.py 2>&1 | tee gpurun_out/tc_diag.log
echo "=== pytest ops"; timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout 600 2>&1 | tail -40 | tee gpurun_out/pytest_ops.log
echo "=== pytest encoder"; timeout 1200 python -m pytest tests/test_gpu_encoder.py -m gpu -q --timeout 900 2>&1 | tail -60 | tee gpurun_out/pytest_enc.log
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench fp32 c1"; timeout 600 python bench.py --dtype fp32 --workload c1 --steps 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_fp32_c1.log
echo "=== bench bf16 c1"; timeout 600 python bench.py --dtype bf16 --workload c1 --steps 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_bf16_c1.log
echo "=== bench bf16 c2"; timeout 600 python bench.py --dtype bf16 --workload c2 --steps 3 --no-cpu-baseline --profile-json gpurun_out/prof_c2.json 2>&1 | tail -3 | tee gpurun_out/bench_bf16_c2.log
echo "=== bench bf16 c3"; timeout 900 python bench.py --dtype bf16 --workload c3 --utts 128 --steps 2 --profile-json gpurun_out/prof_c3.json 2>&1 | tail -3 | tee gpurun_out/bench_bf16_c3.log
