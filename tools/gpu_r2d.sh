#!/bin/bash
mkdir -p gpurun_out
echo "=== super-batch bisect"; timeout 600 python tools/dbg_super.py 2>&1 | tail -80 | tee gpurun_out/dbg_super.log
echo "=== attention + int16 tests"; timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_audio_io.py -m gpu -q -x --timeout 600 -k "attention or int16" 2>&1 | tail -8 | tee gpurun_out/pytest_attn_r2c.log
show='import sys,json
for ln in sys.stdin:
    if not ln.startswith("{"): continue
    d=json.loads(ln); r=d["roofline"]
    print({k:d[k] for k in ("value","ms_per_step","stream_lanes")}, "e2e", d["e2e"]["value"], "frac", r["frac"], d["clocks"]["sm_mhz"])
    print({k:(v["launches"],v["ms"],v["tflops"] or v["gbs"]) for k,v in r["by_kernel"].items() if k.startswith("att") or k.startswith("gemm")})'
for env in "CST_ATTN_V2=1" "CST_ATTN_V2=0"; do
  echo "=== bench c2 $env"; env $env timeout 900 python bench.py --workload c2 --steps 10 --lanes 1 --no-cpu-baseline 2>&1 | tail -2 | python -c "$show"
  echo "=== bench c3 $env"; env $env timeout 900 python bench.py --steps 5 --no-cpu-baseline --profile-json gpurun_out/prof_c3_$env.json 2>&1 | tail -2 | python -c "$show"
done
