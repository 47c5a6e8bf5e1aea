#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu (all)"; ( time timeout 480 python -m pytest tests -m gpu -q --timeout 400 ) 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_all.log
echo "=== dec_rate"; timeout 300 python tools/dec_rate.py --lanes 1 --encode --json gpurun_out/dec_rate.json 2>&1 | tail -4 | cut -c1-900
