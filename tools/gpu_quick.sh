#!/bin/bash
# quick regression + perf check
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_encoder.py -m gpu -q --timeout 900 2>&1 | tail -15 | tee gpurun_out/pytest.log
echo "=== bench bf16 c2"; timeout 600 python bench.py --dtype bf16 --workload c2 --steps 5 --lanes 1 --no-cpu-baseline --profile-json gpurun_out/prof_c2.json 2>&1 | tail -1 | tee gpurun_out/bench_bf16_c2.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','stream_lanes')}, d['e2e']['value'], d['roofline']['by_kernel'])"
echo "=== bench bf16 c3"; timeout 900 python bench.py --dtype bf16 --workload c3 --utts 256 --steps 3 --lanes 3 --no-cpu-baseline --profile-json gpurun_out/prof_c3.json 2>&1 | tail -1 | tee gpurun_out/bench_bf16_c3.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','stream_lanes')}, d['e2e']['value'], d['clocks'], d['roofline']['by_kernel'])"
echo "=== bench bf16 c2 pair=1"; CST_TC_PAIR=1 timeout 600 python bench.py --dtype bf16 --workload c2 --steps 5 --lanes 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','stream_lanes')}, d['e2e']['value'])"
