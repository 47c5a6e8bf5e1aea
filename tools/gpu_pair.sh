#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest ops+encoder"; timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_encoder.py -m gpu -q -x --timeout 600 2>&1 | tail -4
echo "=== rates bulk=3"; timeout 300 python tools/gemm_rate.py 2>&1 | tail -6
echo "=== rates bulk=1"; CST_TC_BULK=1 timeout 300 python tools/gemm_rate.py 2>&1 | tail -6
for l in 3 4 6; do
echo "=== bench bf16 c3 lanes=$l"; timeout 900 python bench.py --dtype bf16 --workload c3 --utts 256 --steps 3 --lanes $l --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','stream_lanes')}, d['e2e']['value'])"
done
echo "=== bench bf16 c2"; timeout 600 python bench.py --dtype bf16 --workload c2 --steps 5 --lanes 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','stream_lanes')}, d['e2e']['value'])"
echo "=== bench bf16 c2 bulk=1"; CST_TC_BULK=1 timeout 600 python bench.py --dtype bf16 --workload c2 --steps 5 --lanes 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','stream_lanes')}, d['e2e']['value'])"
