#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do for pm in 0 1; do
echo "=== c2 pair=$pm"; CST_TC_PAIR=$pm timeout 600 python bench.py --dtype bf16 --workload c2 --steps 10 --lanes 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['clocks']['sm_mhz'], d['roofline']['by_kernel']['gemm_tc_bf16'])"
done; done
for pm in 0 1; do
echo "=== c3 pair=$pm"; CST_TC_PAIR=$pm timeout 900 python bench.py --dtype bf16 --workload c3 --utts 256 --steps 3 --lanes 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['clocks']['sm_mhz'], d['roofline']['by_kernel']['gemm_tc_bf16'])"
done
echo "=== rates pair=1"; CST_TC_PAIR=1 timeout 300 python tools/gemm_rate.py 2>/dev/null
echo "=== rates pair=0"; CST_TC_PAIR=0 timeout 300 python tools/gemm_rate.py 2>/dev/null
