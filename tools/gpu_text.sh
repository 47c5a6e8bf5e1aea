#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest text + plan"; timeout 600 python -m pytest tests/test_text_branch.py -m gpu -x -q --timeout 400 2>&1 | tail -15 | tee gpurun_out/pytest_text.log
bash tools/gpu_dec_ncu.sh
