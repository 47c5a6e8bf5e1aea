#!/bin/bash
# A/B of the 16-byte column-sum kernel on the C5 training step (CST_COLSUM_VEC=0: scalar kernel) + the backward parity tests.
timeout 300 python -m pytest tests/test_gpu_backward.py tests/test_adam.py -x -q 2>&1 | tail -3
for v in 0 1; do
  CST_COLSUM_VEC=$v timeout 120 python bench.py --workload c5 --steps 10 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/ab_colsum_$v.json
  python - "$v" <<'PY'
import sys, json
v = sys.argv[1]
d = json.loads(open("gpurun_out/ab_colsum_%s.json" % v).read())
print("CST_COLSUM_VEC=%s ms_per_step %.3f without_allreduce %.3f colsum %s" % (v, d["ms_per_step"], d["train"]["ms_per_step_without_allreduce"],
                                                                            d["roofline"]["by_kernel"]["colsum"]))
PY
done
