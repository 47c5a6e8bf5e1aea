#!/bin/bash
# end-of-round evidence: full GPU suite, smoke, both bench arms (driver order), c2 bench with per-kernel profile
mkdir -p gpurun_out
echo "=== pytest -m gpu (all)"; timeout 2400 python -m pytest tests -m gpu -q --timeout 1500 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_all.log
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "=== reference arm"; ( time python bench.py --impl reference --steps 3 --warmup 3 ) 2>&1 | tail -5 | tee gpurun_out/bench_reference.log | cut -c1-700
echo "=== default bench"; ( time python bench.py ) 2>&1 | tail -5 | tee gpurun_out/bench_default.log | cut -c1-3000
echo "=== c2 bench"; timeout 600 python bench.py --dtype bf16 --workload c2 --steps 10 --lanes 1 --no-cpu-baseline --profile-json gpurun_out/prof_c2.json 2>&1 | tail -1 | tee gpurun_out/bench_bf16_c2.log | cut -c1-300
