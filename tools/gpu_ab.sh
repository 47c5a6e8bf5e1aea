#!/bin/bash
# A/B template: alternate a run-time switch on the same box (box-to-box spread is +-3 %, larger than most effects)
mkdir -p gpurun_out
for rep in 1 2; do for v in 0 1; do
echo "=== c3 CST_PDL=$v"; CST_PDL=$v timeout 900 python bench.py --utts 256 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['clocks']['sm_mhz'])"
done; done
for rep in 1 2; do for v in 0 1; do
echo "=== c2 CST_PDL=$v"; CST_PDL=$v timeout 600 python bench.py --dtype bf16 --workload c2 --steps 10 --lanes 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['clocks']['sm_mhz'])"
done; done
