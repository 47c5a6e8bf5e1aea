#!/bin/bash
# round 2, second GPU pass: attention v2 bring-up (P in TMEM, persistent), super-batch bit-identity debug
mkdir -p gpurun_out
echo "=== attention op tests (v2 default)"; timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout 600 -k "attention" 2>&1 | tail -15 | tee gpurun_out/pytest_attn_r2b.log
echo "=== super-batch stage diffs"; timeout 600 python tools/dbg_super.py 2>&1 | tail -30 | tee gpurun_out/dbg_super.log
show='import sys,json
for ln in sys.stdin:
    if not ln.startswith("{"): continue
    d=json.loads(ln); r=d["roofline"]
    print({k:d[k] for k in ("value","ms_per_step","stream_lanes")}, "e2e", d["e2e"]["value"], "frac", r["frac"], d.get("super_batches"), d["clocks"]["sm_mhz"])
    print({k:(v["launches"],v["ms"],v["tflops"] or v["gbs"]) for k,v in r["by_kernel"].items()})'
for env in "CST_ATTN_V2=0" "CST_ATTN_V2=1"; do
  echo "=== bench c3 $env"; env $env timeout 900 python bench.py --steps 5 --no-cpu-baseline 2>&1 | tail -2 | python -c "$show"
  echo "=== bench c2 $env"; env $env timeout 900 python bench.py --workload c2 --steps 10 --lanes 1 --no-cpu-baseline 2>&1 | tail -2 | python -c "$show"
done
echo "=== bench c3 98304 rows"; timeout 900 python bench.py --steps 5 --no-cpu-baseline --super-rows 98304 2>&1 | tail -2 | python -c "$show"
echo "=== encoder tests"; timeout 1500 python -m pytest tests/test_gpu_encoder.py tests/test_fairseq_plugin.py -m gpu -q --timeout 1200 2>&1 | tail -15 | tee gpurun_out/pytest_r2b.log
