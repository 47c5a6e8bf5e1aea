#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_encoder.py tests/test_greedy_ids.py -m gpu -q -x --timeout 900 2>&1 | tail -5
echo "=== bench bf16 c3"; timeout 900 python bench.py --dtype bf16 --workload c3 --utts 256 --steps 3 --lanes 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','stream_lanes')}, d['e2e']['value'], d['roofline']['by_kernel']['attention_simt'])"
echo "=== bench bf16 c2"; timeout 600 python bench.py --dtype bf16 --workload c2 --steps 5 --lanes 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','stream_lanes')}, d['e2e']['value'], d['roofline']['by_kernel']['attention_simt'], d['roofline']['by_kernel']['conv0_gn_gelu'])"
