#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2 3; do for pm in 1 3; do
echo "=== c2 pair=$pm"; CST_TC_PAIR=$pm timeout 600 python bench.py --dtype bf16 --workload c2 --steps 10 --lanes 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['clocks']['sm_mhz'], d['roofline']['by_kernel']['gemm_tc_bf16'])"
done; done
for rep in 1 2; do for pm in 0 3; do
echo "=== c3 pair=$pm"; CST_TC_PAIR=$pm timeout 900 python bench.py --utts 256 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['clocks']['sm_mhz'], d['roofline']['by_kernel']['gemm_tc_bf16'])"
done; done
