#!/bin/bash
mkdir -p gpurun_out
echo "=== fp32 encoder tests (split GEMMs on tcgen05)"; timeout 1500 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_loss.py tests/test_text_branch.py -m gpu -q --timeout 1200 2>&1 | tail -8 | cut -c1-300
python - <<'PY'
import sys, torch, time
sys.path.insert(0, '.')
import chimera_st_b200
from chimera_st_b200 import synth
from chimera_st_b200.encoder import build_encoder_from_state_dict
from oracle import chimera_oracle as O
import os
sd = synth.make_state_dict(seed=0, interlingua_length=16)
for lens in ([16000, 12345, 8000], [80000, 64000, 48123, 32000]):
    wave, tl = synth.make_waveforms(lens, seed=7)
    with torch.no_grad():
        ref, _ = O.encoder_forward(sd, wave, tl)
    for mode in ("1", "0"):
        os.environ["CST_F32_TC"] = mode
        enc = build_encoder_from_state_dict(sd, dtype=torch.float32, device="cuda", use_graph=True)
        out = enc(wave.cuda(), tl.cuda()).encoder_out
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): enc(wave.cuda(), tl.cuda())
        e1.record(); torch.cuda.synchronize()
        err = float((out.cpu().double() - ref.double()).norm() / ref.double().norm())
        print("lens", lens, "CST_F32_TC", mode, "rel_l2 %.3e" % err, "ms %.2f" % (e0.elapsed_time(e1) / 5), "audio-s/s %.0f" % (sum(lens) / 16000 / (e0.elapsed_time(e1) / 5e3)))
PY
for env in "CST_F32_TC=1" "CST_F32_TC=0"; do
  echo "=== bench c2 fp32 $env"; env $env timeout 900 python bench.py --workload c2 --dtype fp32 --steps 3 --lanes 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print({k:d[k] for k in ('value','ms_per_step')}, {k:(v['launches'],v['ms']) for k,v in r['by_kernel'].items()})"
done
