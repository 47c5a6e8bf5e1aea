#!/bin/bash
mkdir -p gpurun_out
echo "=== full-size tests"; timeout 1500 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "full_size or margin" --timeout 1200 2>&1 | tail -5
echo "=== default bench (reference arm first, as the driver does)"
( time python bench.py --impl reference --steps 3 --warmup 3 ) 2>&1 | tail -5 | cut -c1-600
( time python bench.py --steps 3 --warmup 3 ) 2>&1 | tail -5 | tee gpurun_out/bench_default.log | cut -c1-2500
