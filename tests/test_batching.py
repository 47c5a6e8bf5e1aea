"""Batching / sharding host logic against goldens produced by the reference's own compiled Cython
(`oracle/build_ref_cython.py` -> tests/golden/batches.npz)."""
import os

import numpy as np
import pytest

import chimera_st_b200  # noqa: F401
from chimera_st_b200 import batching
from conftest import GOLDEN


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
