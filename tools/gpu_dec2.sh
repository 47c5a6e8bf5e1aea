#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest decoder"; timeout 600 python -m pytest tests/test_gpu_decoder.py -m gpu -q --timeout 400 2>&1 | tail -12 | tee gpurun_out/pytest_dec.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4
