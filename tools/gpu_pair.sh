#!/bin/bash
mkdir -p gpurun_out
echo "=== tc_diag"; CST_TC_PAIR=1 timeout 600 python tools/tc_diag.py 2>&1 | tail -24
for pm in 0 2; do
echo "=== pytest ops pair=$pm"; CST_TC_PAIR=$pm timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout 300 2>&1 | tail -4
done
for pm in 0 2; do echo "=== rates pair=$pm"; CST_TC_PAIR=$pm timeout 300 python tools/gemm_rate.py 2>&1 | tail -6; done
echo "=== rates pair=0 bulk=0"; CST_TC_BULK=0 CST_TC_PAIR=0 timeout 300 python tools/gemm_rate.py 2>&1 | tail -6
