#!/bin/bash
mkdir -p gpurun_out
echo "=== loss + encoder tests"; timeout 1500 python -m pytest tests/test_gpu_loss.py tests/test_gpu_encoder.py -m gpu -q --timeout 1200 2>&1 | tail -6 | cut -c1-300
show='import sys,json
for ln in sys.stdin:
    if not ln.startswith("{"): continue
    d=json.loads(ln); r=d["roofline"]
    print({k:d[k] for k in ("value","ms_per_step","stream_lanes")}, "e2e", d["e2e"]["value"], "frac", r["frac"], d["clocks"]["sm_mhz"], d.get("parity") and d["parity"]["worst_row_rel_l2"])
    print({k:(v["launches"],v["ms"],v["tflops"] or v["gbs"]) for k,v in r["by_kernel"].items() if k[:4] in ("gemm","laye","atte","dec_")})'
for env in "CST_LN_FUSE=1" "CST_LN_FUSE=0" "CST_LN_FUSE=1 CST_MEM_FUSED=0"; do
  echo "=== bench c2 $env"; env $env timeout 900 python bench.py --workload c2 --steps 10 --lanes 1 --no-cpu-baseline 2>&1 | tail -2 | python -c "$show"
  echo "=== bench c3 $env"; env $env timeout 900 python bench.py --steps 5 --no-cpu-baseline 2>&1 | tail -2 | python -c "$show"
done
echo "=== bench c3 49152 with parity"; timeout 900 python bench.py --steps 5 --super-rows 49152 2>&1 | tail -2 | python -c "$show"
