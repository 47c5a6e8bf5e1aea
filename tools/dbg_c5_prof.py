import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import chimera_st_b200  # noqa
from chimera_st_b200 import synth, losses
from chimera_st_b200.train import EncoderTrainStep
import bench
B, Lw, M = 8, 150000, 16
sd = synth.make_state_dict(seed=0, interlingua_length=16, dead_heads=False)
wave, tl = synth.make_waveforms([Lw] * B, seed=3)
wave, tl = wave.cuda(), tl.cuda()
step = EncoderTrainStep(sd, B, Lw, device="cuda", dtype=torch.bfloat16)
dmem = torch.randn(M, B, 512).cuda()
class Trace:
    def __init__(self, lib): self.lib = lib; self.n = 0
    def __getattr__(self, name):
        fn = getattr(self.lib, name)
        if not name.startswith("cst_") or name in ("cst_last_error",): return fn
        def t(*a):
            rc = fn(*a)
            try:
                torch.cuda.synchronize()
            except Exception as e:
                print("FAULT after call #%d %s args=%s" % (self.n, name, [x for x in a if isinstance(x, int)][:30])); sys.stdout.flush(); raise
            self.n += 1
            return rc
        return t
mode = sys.argv[1] if len(sys.argv) > 1 else "trace"
if mode == "trace":
    step.o.lib = Trace(step.o.lib)
else:
    step.o.lib = bench.LaunchProfiler(step.o.lib)
step.forward(wave, tl); step.backward(dmem); torch.cuda.synchronize()
print("ok", mode)
