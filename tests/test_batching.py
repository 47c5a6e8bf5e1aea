"""Batching / sharding host logic against goldens produced by the reference's own compiled Cython
(`oracle/build_ref_cython.py` -> tests/golden/batches.npz)."""
import os

import numpy as np
import pytest

import chimera_st_b200  # noqa: F401
from chimera_st_b200 import batching
from conftest import GOLDEN, ROOT


@pytest.mark.parametrize("name", ["c3", "small", "maxsent", "mult1"])
def test_batch_by_size_matches_reference_cython(name):
    g = np.load(os.path.join(GOLDEN, "batches.npz"))
    lens, flat, sizes = g[name + "_lens"], g[name + "_flat"], g[name + "_sizes"]
    mt, ms, mult = g[name + "_cfg"].tolist()
    got = batching.batch_by_size(batching.ordered_indices(lens), lens, mt, ms, mult)
    assert [len(b) for b in got] == sizes.tolist()
    assert [i for b in got for i in b] == flat.tolist()
    for b in got:                      # the token budget holds for every batch
        assert len(b) * max(int(lens[i]) for i in b) <= mt


def test_shards_partition_the_batches_round_robin():
    batches = [[i] for i in range(11)]
    shards = [batching.shard_batches(batches, 4, r) for r in range(4)]
    assert all(len(s) == 3 for s in shards)                       # padded to equal length
    assert sorted(i for s in shards for b in s for i in b) == list(range(11))
    assert shards[1][:3] == [[1], [5], [9]] and shards[3][-1] == []
    with pytest.raises(ValueError):
        batching.shard_batches(batches, 4, 4)


def test_collate_matches_reference_collater_rule():
    """zero padding to the longest + descending-length order (torch sort), ids follow the rows"""
    import torch
    from chimera_st_b200 import batching as Bt
    g = torch.Generator().manual_seed(0)
    lens = [5, 9, 9, 1, 7]
    waves = [torch.randn(n, generator=g) for n in lens]
    ids, x, n = Bt.collate_waveforms(waves, ids=[10, 11, 12, 13, 14])
    assert x.shape == (5, 9) and n.tolist() == sorted(lens, reverse=True)
    exp_n, order = torch.tensor(lens).sort(descending=True)
    assert ids.tolist() == [[10, 11, 12, 13, 14][i] for i in order.tolist()]
    for row, i in enumerate(order.tolist()):
        assert torch.equal(x[row, :lens[i]], waves[i]) and not bool(x[row, lens[i]:].any())


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not mounted")
def test_collate_against_the_reference_function():
    import subprocess
    import sys
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from oracle import make_overlay
make_overlay.build(); make_overlay.activate()
import fairseq.models
from fairseq.data.audio.speech_to_text_dataset import _collate_frames
import chimera_st_b200
from chimera_st_b200 import batching as Bt
g = torch.Generator().manual_seed(3)
lens = [1200, 800, 1200, 5, 977, 800]
waves = [torch.randn(n, generator=g) for n in lens]
frames = _collate_frames(waves, True)
n = torch.tensor(lens)
n, order = n.sort(descending=True)                      # triplet_dataset.py:171-179
frames = frames.index_select(0, order)
ids, x, m = Bt.collate_waveforms(waves)
assert torch.equal(x, frames) and torch.equal(m, n) and torch.equal(ids, order)
print("COLLATE_OK")
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    assert "COLLATE_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


def test_encode_utterances_routes_every_utterance_once():
    """host logic of the feeder with a stand-in encoder: reference batching, sharding, id bookkeeping"""
    import torch
    from chimera_st_b200 import batching as Bt

    class Out:
        def __init__(self, t):
            self.encoder_out = t

    class FakeEncoder:
        def forward_many(self, batches, n_lanes=3):
            outs = []
            for w, n in batches:
                assert int(n.max()) == w.shape[1] and n.tolist() == sorted(n.tolist(), reverse=True)
                assert w.shape[0] * w.shape[1] <= 40000 or w.shape[0] == 1
                # "memories": M = 2 rows of [sum of the samples, valid length] per utterance
                s = torch.stack([w.sum(1), n.float()], 1)                       # [B, 2]
                outs.append(Out(s.t().unsqueeze(-1).expand(2, w.shape[0], 512).clone()))
            return outs
    g = torch.Generator().manual_seed(1)
    lens = torch.randint(100, 9000, (37,), generator=g).tolist()
    waves = [torch.randn(n, generator=g) for n in lens]
    got = {}
    for shard in range(2):
        part = Bt.encode_utterances(FakeEncoder(), waves, max_tokens=40000, bsz_mult=8, num_shards=2, shard_id=shard)
        assert not (set(part) & set(got))
        got.update(part)
    assert sorted(got) == list(range(37))
    for i, w in enumerate(waves):
        assert abs(float(got[i][0, 0]) - float(w.sum())) < 1e-2 and int(got[i][1, 0]) == lens[i]
