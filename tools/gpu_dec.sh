#!/bin/bash
# decoder bring-up: kernel + hypothesis parity, then decode-step latency at the C4 shape (plain and with PDL)
mkdir -p gpurun_out
echo "=== pytest decoder"; timeout 900 python -m pytest tests/test_gpu_decoder.py -m gpu -x -q --timeout 600 2>&1 | tail -25 | tee gpurun_out/pytest_dec.log
echo "=== dec_rate"; timeout 600 python tools/dec_rate.py --encode --json gpurun_out/dec_rate.json 2>&1 | tail -8 | cut -c1-1500
echo "=== dec_rate no PDL"; CST_DEC_PDL=0 timeout 600 python tools/dec_rate.py --lanes 1,4 --json gpurun_out/dec_rate_pdl.json 2>&1 | tail -8 | cut -c1-400
