#!/bin/bash
# memcheck + racecheck of the decoder / text-branch kernels (tiny shapes, eager launches)
mkdir -p gpurun_out
if [ "$1" != "race" ]; then
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_dec.py 2>&1 | grep -v "^$" | tail -12 > gpurun_out/sanitizer_memcheck_dec.log; cat gpurun_out/sanitizer_memcheck_dec.log
fi
SANITIZE_ONLY=bf16 timeout 120 compute-sanitizer --tool racecheck python tools/sanitize_dec.py 2>&1 | grep -v "^$" | grep -E "Race reported|Read access|RACECHECK|decode|text|ERROR" | sort | uniq -c | sort -rn | head -30 > gpurun_out/sanitizer_racecheck_dec.log; cat gpurun_out/sanitizer_racecheck_dec.log
