"""CPU: smallest |pre-activation| of each of the nine ReLU layers of the oracle forward (fp64) for a few seeds -- there is always one within
~1e-6 of zero, which is why tests/test_gpu_backward.py pins the ReLU sign pattern when it compares gradients."""
import sys; sys.path.insert(0,'/root/repo')
import torch
import chimera_st_b200
from chimera_st_b200 import synth
from oracle import chimera_oracle as O
torch.set_num_threads(8)
sd = synth.make_state_dict(seed=0, interlingua_length=16, dead_heads=False)
sd64 = {k:(v.double() if v.is_floating_point() else v) for k,v in sd.items()}
orig = torch.relu
for lens in ([6000,4500],[16000,12345,8000]):
    for seed in range(31,40):
        wave, tl = synth.make_waveforms(lens, seed=seed)
        mins=[]
        def rec(z):
            mins.append(float(z.abs().min())); return orig(z)
        torch.relu = rec
        with torch.no_grad():
            O.encoder_forward(sd64, wave.double(), tl)
        torch.relu = orig
        print(lens, seed, "min|z| over %d relu layers: %.2e"%(len(mins), min(mins)), ["%.1e"%m for m in mins])
