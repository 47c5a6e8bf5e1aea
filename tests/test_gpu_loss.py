"""Training heads on the GPU against the oracle restatements of the reference criteria, values and
gradients (autograd through the oracle)."""
import pytest
import torch

import chimera_st_b200  # noqa: F401
from chimera_st_b200 import losses
from oracle import loss_oracle as LO
from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,B", [(16, 5), (64, 3), (1, 2)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-5)])
def test_contrastive_loss_and_gradients(M, B, dtype, tol):
    g = torch.Generator().manual_seed(M + B)
    a = torch.randn(M, B, 512, generator=g)
    t = 0.6 * a + 0.4 * torch.randn(M, B, 512, generator=g)
    a, t = a.to(dtype), t.to(dtype)
    ar, tr = a.float().requires_grad_(), t.float().requires_grad_()       # bf16 inputs: the kernel reads them exactly
    ref_rows = LO.contrastive(ar, tr, 0.1, reduce=False)
    ref = ref_rows.sum()
    (0.37 * ref).backward()
    rows, total, da, dt = losses.contrastive_loss(a.cuda(), t.cuda(), temp=0.1, grad_scale=0.37)
    assert rows.shape == (B, M)
    # loss = lse - logit with logits ~ 1/temp = 10: an fp32 cancellation on both sides, so compare on the logit scale
    assert float((rows.cpu() - ref_rows.detach()).abs().max()) < 1e-5
    assert abs(float(total) - float(ref.detach())) <= 1e-5 * rows.numel()
    # (softmax - 1 at the target cancels to ~3e-3 when the diagonal dominates: fp32 on both sides)
    if M > 1:            # (M == 1: the softmax has one class, loss and gradients are exactly zero)
        assert rel_l2(da.cpu(), ar.grad) < 5e-4 and rel_l2(dt.cpu(), tr.grad) < 5e-4
    else:
        assert float(da.abs().max()) < 1e-6 and float(dt.abs().max()) < 1e-6
    rows2, total2, _, _ = losses.contrastive_loss(a.cuda(), t.cuda(), temp=0.1, grad_scale=None)
    assert torch.equal(rows2, rows) and _ is None


@pytest.mark.parametrize("N,V", [(37, 10000), (5, 777), (300, 10000)])
def test_label_smoothed_ce_and_gradient(N, V):
    g = torch.Generator().manual_seed(N)
    logits = (torch.randn(N, V, generator=g) * 2.5).requires_grad_()
    tg = torch.randint(0, V, (N,), generator=g)
    tg[::7] = 1                                                            # padding positions
    lp = torch.log_softmax(logits, -1)
    loss_r, nll_r = LO.label_smoothed_nll(lp, tg, 0.1, ignore_index=1, reduce=False)
    (loss_r.sum() * 0.5).backward()
    out = losses.label_smoothed_ce(logits.detach().cuda(), tg.cuda(), eps=0.1, ignore_index=1, grad_scale=0.5)
    assert rel_l2(out["loss_rows"].cpu(), loss_r.detach().reshape(-1)) < 2e-6
    assert rel_l2(out["nll_rows"].cpu(), nll_r.detach().reshape(-1)) < 2e-6
    assert abs(float(out["loss"]) - float(loss_r.sum())) < 2e-5 * float(loss_r.sum())
    assert rel_l2(out["dlogits"].cpu(), logits.grad) < 2e-5
    assert not bool(out["dlogits"][::7].any())
