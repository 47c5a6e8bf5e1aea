#!/bin/bash
mkdir -p gpurun_out
echo "=== loss tests"; timeout 600 python -m pytest tests/test_gpu_loss.py -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
echo "=== fused LN op tests"; timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 600 -k "fused_layernorm" 2>&1 | tail -4 | cut -c1-300
for env in "CST_LN_FUSE=1" "CST_LN_FUSE=0"; do
  echo "=== bench c3 $env"; env $env timeout 900 python bench.py --steps 3 --no-cpu-baseline --profile-json gpurun_out/prof_c3_$env.json 2>&1 | tail -1 | cut -c1-200
done
