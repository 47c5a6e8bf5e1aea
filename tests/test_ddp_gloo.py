"""Gradient all-reduce plumbing of the training configuration (C5) on CPU: world_size-2 gloo, bucketed + asynchronous, against
the reference's rule (flat buffer, pre-divided by the world size, summed: legacy_distributed_data_parallel.py:94-178)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import chimera_st_b200  # noqa: F401
from chimera_st_b200 import ddp

SIZES = [("mem.fc2.weight", 2048 * 16), ("mem.fc2.bias", 512), ("big.weight", 40000), ("enc.ln.weight", 512), ("enc.qkv.weight", 9000),
         ("missing.on.rank1", 77), ("conv0.weight", 5120)]


def _worker(rank, world, port, q):
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    grads = {n: torch.randn(sz, generator=g) for n, sz in SIZES}
    if rank == 1:
        del grads["missing.on.rank1"]                      # a parameter without a gradient on one rank counts as zeros
    red = ddp.GradAllReducer(SIZES, bucket_bytes=64 * 1024)
    # gradients arrive in backward order, a few at a time: complete buckets start reducing before the rest exists
    names = [n for n, _ in SIZES if n in grads]
    red.ready({n: grads[n] for n in names[:3]})
    started_early = len(red.inflight)
    red.ready({n: grads[n] for n in names[3:]})
    out = red.finish(grads)
    q.put((rank, started_early, len(red.buckets), {n: t.numpy().copy() for n, t in out.items()}))   # plain arrays: no fd passing
    dist.destroy_process_group()


def test_bucket_plan_keeps_backward_order_and_isolates_big_tensors():
    b = ddp.plan_buckets(SIZES, bucket_bytes=64 * 1024)
    assert [n for bk in b for n in bk] == [n for n, _ in SIZES]
    assert ["mem.fc2.weight"] in b and ["big.weight"] in b            # >= one bucket (16384 elements): travel alone
    assert all(sum(dict(SIZES)[n] for n in bk) <= 16384 or len(bk) == 1 for bk in b)


def test_two_rank_gloo_all_reduce_matches_the_reference_rule():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, early0, nb, out0), (_, early1, _, out1) = res
    assert early0 >= 1 and early1 >= 1 and nb >= 3                   # overlap: something was in flight before the last gradient
    gens = [torch.Generator().manual_seed(100 + r) for r in range(2)]
    ref = {n: [torch.randn(sz, generator=gens[r]) for r in range(2)] for n, sz in SIZES}
    for n, sz in SIZES:
        want = ref[n][0] / 2 + (ref[n][1] / 2 if n != "missing.on.rank1" else 0)
        assert torch.allclose(torch.from_numpy(out0[n]), want, atol=1e-7) and (out0[n] == out1[n]).all(), n


def test_persistent_buffers_keep_their_addresses_and_bf16_wire_format():
    """GradAllReducer(persistent=True): the flat buffers (and therefore the reduced-gradient views a CUDA-graphed optimizer update reads)
    are allocated once; comm_dtype=bfloat16 rounds the pre-divided gradients to the wire format (single process: no collective)."""
    sizes = [("a", 300), ("b", 40000), ("c", 7)]
    red = ddp.GradAllReducer(sizes, world_size=1, bucket_bytes=64 * 1024, comm_dtype=torch.bfloat16, persistent=True)
    g = torch.Generator().manual_seed(3)
    ptrs = None
    for step in range(3):
        grads = {n: torch.randn(sz, generator=g) for n, sz in sizes}
        red.ready(grads)
        out = red.finish(grads)
        for n, sz in sizes:
            assert out[n].dtype == torch.bfloat16 and out[n].numel() == sz
            assert torch.equal(out[n].float(), grads[n].to(torch.bfloat16).float())
        now = {n: out[n].data_ptr() for n, _ in sizes}
        assert ptrs is None or now == ptrs
        ptrs = now


def test_layerdrop_draw_follows_the_reference_rule():
    """One uniform number per wav2vec2 layer, the layer runs iff the number exceeds p (wav2vec2.py:835-838)."""
    import numpy as np
    from chimera_st_b200.train import EncoderTrainStep
    for seed, p in ((0, 0.3), (5, 0.05), (9, 0.9)):
        u = np.random.RandomState(seed).random_sample(12)
        want = frozenset(i for i in range(12) if not u[i] > p)
        assert EncoderTrainStep.sample_layerdrop(p, np.random.RandomState(seed)) == want
    assert EncoderTrainStep.sample_layerdrop(0.0, np.random.RandomState(1)) == frozenset()


_PROBE = ("interlingua_layers.2.fc2.weight", "transformer_layers.0.self_attn.q_proj.weight", "wav2vec_model.encoder.layers.11.fc1.bias",
          "wav2vec_model.encoder.layers.3.self_attn.out_proj.weight", "wav2vec_model.post_extract_proj.weight",
          "wav2vec_model.feature_extractor.conv_layers.0.0.weight")


def _rank_inputs(rank):
    from chimera_st_b200 import synth
    sd = synth.make_state_dict(seed=0, interlingua_length=8, dead_heads=False)
    wave, tl = synth.make_waveforms([3000, 2400], seed=40 + rank)            # same shapes, different audio per rank (data parallel)
    R = torch.randn(8, 2, 512, generator=torch.Generator().manual_seed(5))
    return sd, wave, tl, R


def _train_worker(rank, world, port, q):
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emu import EmuLib
    from chimera_st_b200.train import EncoderTrainStep
    torch.set_num_threads(4)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sd, wave, tl, R = _rank_inputs(rank)
    step = EncoderTrainStep(sd, 2, wave.shape[1], device="cpu", lib=EmuLib())
    step.forward(wave, tl)
    parts = list(step.backward_iter(R))                             # the four backward segments, in the order they finish
    names = [(k, v.numel()) for part in parts for k, v in part.items()]                        # backward order = bucket order
    red = ddp.GradAllReducer(names, bucket_bytes=8 << 20)
    grads, early = {}, []
    for part in parts:                                              # hand-over per segment: finished buckets reduce before the rest arrives
        red.ready(part)
        grads.update(part)
        early.append(len(red.inflight))
    out = red.finish(grads)
    q.put((rank, early, len(red.buckets), len(names), {n: out[n].reshape(-1)[:4096].numpy().copy() for n in _PROBE},
           float(sum(float(v.double().sum()) for v in out.values()))))
    dist.destroy_process_group()


def test_two_rank_emulated_training_step_with_overlapped_all_reduce():
    """The N > 1 training path end to end on CPU: two gloo ranks run the real `EncoderTrainStep` launch sequence (ABI emulator) on
    different audio, hand the gradients of each backward segment to `GradAllReducer` as they finish, and end with the average of the two
    ranks' gradients (legacy_distributed_data_parallel.py:94-178: pre-divide by the world size, sum), identical on both ranks."""
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emu import EmuLib
    from chimera_st_b200.train import EncoderTrainStep
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    # the expected result, serially, while the ranks work
    want = None
    for r in range(2):
        sd, wave, tl, R = _rank_inputs(r)
        _, G = EncoderTrainStep(sd, 2, wave.shape[1], device="cpu", lib=EmuLib()).forward_backward(wave, tl, R)
        want = {k: v / 2 for k, v in G.items()} if want is None else {k: want[k] + G[k] / 2 for k in want}
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, early0, nb, nn, probe0, tot0), (_, early1, _, _, probe1, tot1) = res
    assert nn == 361 and nb >= 8
    assert early0[0] >= 1 and early0 == sorted(early0) and early1[0] >= 1    # buckets in flight after the FIRST segment already
    assert tot0 == tot1                                                      # bit-identical on both ranks
    for n in _PROBE:
        assert (probe0[n] == probe1[n]).all(), n
        ref = want[n].reshape(-1)[:4096]
        err = float((torch.from_numpy(probe0[n]) - ref).norm() / ref.norm())
        assert err < 1e-5, (n, err)
