"""The loss restatements (oracle/loss_oracle.py) against the UNMODIFIED reference criterion code (mounted tree or the shipped
bundle)."""
import subprocess
import sys
import os

import pytest

from conftest import ROOT
from oracle import make_overlay

pytestmark = pytest.mark.skipif(not make_overlay.available(), reason="no reference tree (mount or oracle/_ref/src bundle)")

SCRIPT = r'''
import sys, torch
sys.path.insert(0, %(root)r)
from oracle import make_overlay, loss_oracle as LO
make_overlay.build(); make_overlay.activate()
import fairseq.models, fairseq.criterions
from fairseq.criterions.label_smoothed_cross_entropy import label_smoothed_nll_loss
from fairseq.criterions.triplet_st_mt_contrastive import TripletSTMTContrastiveCriterion as Crit
g = torch.Generator().manual_seed(0)
for M, B in ((16, 3), (64, 2)):
    a, t = torch.randn(M, B, 512, generator=g), torch.randn(M, B, 512, generator=g)
    t = 0.7 * a + 0.3 * t
    class Fake: contrastive_temp = 0.1
    for reduce in (True, False):
        ref = Crit.compute_contrastive(Fake(), a, t, reduce)
        got = LO.contrastive(a, t, 0.1, reduce)
        assert torch.allclose(ref.reshape(-1), got.reshape(-1), rtol=1e-6, atol=1e-6), (ref, got)
lp = torch.log_softmax(torch.randn(37, 1000, generator=g) * 3, -1)
tg = torch.randint(0, 1000, (37,), generator=g); tg[5] = 1; tg[20] = 1
for reduce in (True, False):
    r = label_smoothed_nll_loss(lp.clone(), tg, 0.1, ignore_index=1, reduce=reduce)
    o = LO.label_smoothed_nll(lp, tg, 0.1, ignore_index=1, reduce=reduce)
    assert torch.allclose(r[0], o[0]) and torch.allclose(r[1], o[1])
print("LOSS_ORACLE_OK")
'''


def test_loss_oracle_matches_reference_criteria():
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=600, env=env)
    assert "LOSS_ORACLE_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
