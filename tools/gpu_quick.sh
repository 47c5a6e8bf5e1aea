#!/bin/bash
# quick regression + perf check
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_encoder.py -m gpu -q --timeout 900 2>&1 | tail -25 | tee gpurun_out/pytest.log
echo "=== bf16 err"; timeout 600 python tools/bf16_err.py 2>&1 | tail -3 | tee gpurun_out/bf16_err.log
echo "=== bench bf16 c2"; timeout 600 python bench.py --dtype bf16 --workload c2 --steps 5 --no-cpu-baseline --profile-json gpurun_out/prof_c2.json 2>&1 | tail -1 | tee gpurun_out/bench_bf16_c2.log
echo "=== bench bf16 c3"; timeout 900 python bench.py --dtype bf16 --workload c3 --utts 256 --steps 2 --no-cpu-baseline --profile-json gpurun_out/prof_c3.json 2>&1 | tail -1 | tee gpurun_out/bench_bf16_c3.log
