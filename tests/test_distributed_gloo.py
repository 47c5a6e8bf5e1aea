"""N>1 host path on CPU: world_size-2 gloo processes shard a global utterance list with the product's
sharding code, each 'encodes' its batches (oracle stand-in for the kernels: the point here is the
plumbing), and the gathered result must equal the single-process result batch for batch; the scalar
reductions used by bench.py (sum of audio seconds, max of time) are checked too."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

LENS = [5200, 4100, 6400, 3300, 4800, 3900, 6000, 3500, 4400, 5600, 3700]


def _encode(batch_lens, sd):
    from chimera_st_b200 import synth
    from oracle import chimera_oracle as O
    wave, lens = synth.make_waveforms(batch_lens, seed=sum(batch_lens) % 9973)
    with torch.no_grad():
        feats = O.conv_feature_extractor(sd, wave)           # cheap stage: enough to fingerprint a batch
    return float(feats.double().sum())


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import chimera_st_b200  # noqa: F401
    from chimera_st_b200 import distributed as D, synth
    r, w, _ = D.init("gloo")
    assert (r, w) == (rank, world)
    sd = synth.make_state_dict(seed=0, dead_heads=False)
    lens = np.asarray(LENS)
    mine, n_total = D.shard_utterances(lens, w, r, max_tokens=12000, bsz_mult=2)
    res = {tuple(b): _encode([int(lens[i]) for i in b], sd) for b in mine}
    gathered = [None] * w
    dist.all_gather_object(gathered, res)
    audio = D.reduce_sum(sum(int(lens[i]) for b in mine for i in b) / 16000.0)
    tmax = D.reduce_max(1.0 + rank)
    D.barrier()
    if rank == 0:
        q.put((gathered, n_total, audio, tmax))
    D.finalize()


def test_two_rank_sharding_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, n_total, audio, tmax = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import chimera_st_b200  # noqa: F401
    from chimera_st_b200 import distributed as D, synth
    lens = np.asarray(LENS)
    single, n1 = D.shard_utterances(lens, 1, 0, max_tokens=12000, bsz_mult=2)
    assert n1 == n_total == len(single)
    merged = {}
    for g in gathered:
        assert not (set(g) & set(merged))                 # shards are disjoint
        merged.update(g)
    assert set(merged) == {tuple(b) for b in single}      # ... and cover every batch exactly once
    sd = synth.make_state_dict(seed=0, dead_heads=False)
    for b in single:                                      # rank-sharded result == single-process result
        assert merged[tuple(b)] == _encode([int(lens[i]) for i in b], sd)
    assert abs(audio - sum(LENS) / 16000.0) < 1e-9 and tmax == 2.0
    assert [tuple(b) for b in single[0::2]] == list(gathered[0]) and [tuple(b) for b in single[1::2]] == list(gathered[1])
