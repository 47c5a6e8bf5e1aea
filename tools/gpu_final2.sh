#!/bin/bash
# end-of-session evidence: full GPU suite, smoke, default bench, c4 encode+decode bench
mkdir -p gpurun_out
echo "=== pytest -m gpu (all)"; ( time timeout 480 python -m pytest tests -m gpu -q -x --timeout 400 ) 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_all.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "=== default bench"; ( time timeout 400 python bench.py --no-cpu-baseline ) 2>&1 | tail -5 | tee gpurun_out/bench_default.log | cut -c1-1500
echo "=== c4 bench + decode"; timeout 300 python bench.py --workload c4 --lanes 1 --steps 5 --no-cpu-baseline --decode 2>&1 | tail -1 | tee gpurun_out/bench_c4_decode.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['decode'])"
