#!/bin/bash
# backward parity tests + one timing of the C5 step (with and without the dropout recipe)
timeout 300 python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | tail -3
for dr in 0 0.1; do
  timeout 120 python bench.py --workload c5 --dropout $dr --steps 10 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/c5_quick_$dr.json
  python - "$dr" <<'PY'
import sys, json
v = sys.argv[1]
d = json.loads(open("gpurun_out/c5_quick_%s.json" % v).read())
print("dropout=%s ms_per_step %.3f without_allreduce %.3f" % (v, d["ms_per_step"], d["train"]["ms_per_step_without_allreduce"]))
PY
done
