#!/bin/bash
# round 2, first GPU pass: new tests (super-batch, every-row parity, plugin under real fairseq), then the bench sweep
mkdir -p gpurun_out
echo "=== pytest (new + encoder)"; timeout 1500 python -m pytest tests/test_gpu_encoder.py tests/test_fairseq_plugin.py tests/test_gpu_beam.py -m gpu -q -x --timeout 1200 2>&1 | tail -25 | tee gpurun_out/pytest_r2a.log
show='import sys,json
for ln in sys.stdin:
    if not ln.startswith("{"): continue
    d=json.loads(ln); r=d["roofline"]
    print({k:d[k] for k in ("value","ms_per_step","stream_lanes")}, "e2e", d["e2e"]["value"], "frac", r["frac"], d.get("super_batches"), d["clocks"]["sm_mhz"])
    print({k:(v["launches"],v["ms"],v["tflops"] or v["gbs"]) for k,v in r["by_kernel"].items()})
    print("parity", d.get("parity"), "cpu", d.get("cpu_baseline"))'
for cfg in "--super-rows 0 --lanes 3" "--lanes 1" "--lanes 2" "--lanes 3" "--super-rows 49152 --lanes 2" "--super-rows 12288 --lanes 2"; do
  echo "=== bench c3 $cfg"; timeout 900 python bench.py --steps 5 --no-cpu-baseline $cfg 2>&1 | tail -2 | tee -a gpurun_out/bench_r2a.log | python -c "$show"
done
echo "=== bench default (with cpu baseline + parity)"; timeout 900 python bench.py --steps 5 2>&1 | tail -2 | tee gpurun_out/bench_default_r2a.log | python -c "$show"
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 10 --warmup 2 2>&1 | tail -1 | tee gpurun_out/bench_ref_r2a.log | cut -c1-600
