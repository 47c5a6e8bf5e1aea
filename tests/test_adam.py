"""Optimizer of the training configuration: the oracle restatement of fairseq/optim/adam.py against goldens of
the unmodified class (CPU), and the fused CUDA kernel behind `train.FusedAdam` against the same goldens (GPU)."""
import os

import numpy as np
import pytest
import torch

import chimera_st_b200  # noqa: F401
from oracle import adam_oracle
from conftest import GOLDEN, rel_l2


def _gold():
    g = np.load(os.path.join(GOLDEN, "adam.npz"))
    hp = dict(lr=float(g["lr"]), betas=tuple(float(x) for x in g["betas"]), eps=float(g["eps"]), weight_decay=float(g["weight_decay"]))
    return g, hp


def test_adam_oracle_matches_the_reference_optimizer():
    g, hp = _gold()
    p = torch.from_numpy(g["p0"]).clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for t in range(3):
        adam_oracle.adam_step(p, torch.from_numpy(g["grads"][t]), m, v, t + 1, **hp)
        assert torch.equal(p, torch.from_numpy(g["after"][t]))           # same torch expressions: bit-exact


@pytest.mark.gpu
@pytest.mark.parametrize("gdtype", [torch.float32, torch.bfloat16])
def test_fused_adam_kernel_matches_the_reference_optimizer(gdtype):
    from chimera_st_b200.train import FusedAdam
    g, hp = _gold()
    params = {"w": torch.from_numpy(g["p0"]).clone().cuda()}
    opt = FusedAdam(params, **hp)
    ref = torch.from_numpy(g["p0"]).clone()
    m, v = torch.zeros_like(ref), torch.zeros_like(ref)
    for t in range(3):
        gr = torch.from_numpy(g["grads"][t]).to(gdtype)
        opt.advance()
        opt.step({"w": gr.cuda().contiguous()})
        adam_oracle.adam_step(ref, gr.float(), m, v, t + 1, **hp)          # the reference on the SAME (possibly bf16-rounded) gradient
        assert rel_l2(params["w"].cpu(), ref) < 2e-6
        if gdtype == torch.float32:
            assert rel_l2(params["w"].cpu(), torch.from_numpy(g["after"][t])) < 2e-6
    # grad_scale (clipping coefficient / inverse loss scale) and a parameter without a gradient
    p2 = {"a": torch.ones(5, device="cuda"), "b": torch.ones(5, device="cuda")}
    o2 = FusedAdam(p2, lr=0.1, betas=(0.0, 0.0), eps=0.0)
    o2.advance(grad_scale=0.0)
    o2.step({"a": torch.ones(5, device="cuda")})
    assert torch.equal(p2["b"].cpu(), torch.ones(5))
    assert not torch.isfinite(p2["a"]).all() or True                         # 0 / 0 with eps = 0 is the caller's problem; no crash


def test_inverse_sqrt_schedule_matches_the_reference_class():
    """train.InverseSqrtLR against goldens of the unmodified InverseSquareRootSchedule (oracle/gen_golden_lr.py): bit-identical floats."""
    from chimera_st_b200.train import InverseSqrtLR
    g = np.load(os.path.join(GOLDEN, "lr_schedule.npz"))
    for name in ("recipe", "init"):
        lr, warm, init = (float(x) for x in g[name + "_cfg"])
        sch = InverseSqrtLR(lr, int(warm), init)
        assert sch.initial == float(g[name + "_initial"])
        got = [sch.at(int(u)) for u in g["updates"]]
        assert got == [float(x) for x in g[name + "_lr"]], (got, g[name + "_lr"])
