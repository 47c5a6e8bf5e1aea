"""End-to-end GPU parity of the B200 encoder against the golden vectors of the UNMODIFIED reference
(tests/golden, fp32) and against the oracle on fresh seeded inputs.
Tolerances (BASELINE.json north_star): fp32 <= 1e-5, bf16 <= 1e-2 relative (rel-L2 per tensor against
the fp32 reference on the identical padded batch); masks / lengths bit-exact."""
import os

import numpy as np
import pytest
import torch

import chimera_st_b200  # noqa: F401
from chimera_st_b200 import synth
from oracle import chimera_oracle as O
from conftest import rel_l2, rel_max, GOLDEN

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-5, torch.bfloat16: 1e-2}
_cache = {}


def encoder(M, dtype, use_graph=False):
    from chimera_st_b200.encoder import build_encoder_from_state_dict
    key = (M, dtype, use_graph)
    if key not in _cache:
        sd = synth.make_state_dict(seed=0, interlingua_length=M)
        _cache[key] = build_encoder_from_state_dict(sd, dtype=dtype, device="cuda", use_graph=use_graph)
    return _cache[key]


def _golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    wave, lens = synth.make_waveforms(g["src_lengths"].tolist(), seed=int(g["wave_seed"]))
    return g, wave, lens


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_tiny_all_stages(dtype):
    g, wave, lens = _golden("tiny")
    enc = encoder(16, dtype)
    feat, fmask, flen = enc._get_w2v_feature(wave.cuda(), lens.cuda())
    assert torch.equal(fmask.cpu(), torch.from_numpy(g["frame_mask"]))
    assert torch.equal(flen.cpu(), torch.from_numpy(g["w2v_len"]))
    assert feat.shape == (3, 49, 768)
    tol = TOL[dtype]
    assert rel_l2(feat.cpu(), torch.from_numpy(g["w2v_out"])) < tol
    out = enc(wave.cuda(), lens.cuda(), mask=None)          # tolerate the collater's stray kwarg
    plan = enc._plan(*wave.shape)
    assert rel_l2(plan.view("conv_feats").cpu(), torch.from_numpy(g["conv_feats"])) < tol
    assert rel_l2(plan.view("h_enc").cpu(), torch.from_numpy(g["h_enc"])) < tol
    assert out.encoder_out.shape == (16, 3, 512) and out.encoder_out.dtype == torch.float32
    assert rel_l2(out.encoder_out.cpu(), torch.from_numpy(g["memories"])) < tol
    assert rel_max(out.encoder_out.cpu(), torch.from_numpy(g["memories"])) < 10 * tol
    assert torch.equal(out.encoder_padding_mask.cpu(), torch.from_numpy(g["encoder_padding_mask"]))
    assert plan.sub_valid.tolist() == g["sub_len"].tolist()
    assert out.encoder_embedding is None and out.encoder_states is None


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", ["c1", "c1mix"])
def test_c1_memories(name, dtype):
    g, wave, lens = _golden(name)
    enc = encoder(16, dtype)
    out = enc(wave.cuda(), lens.cuda())
    assert rel_l2(out.encoder_out.cpu(), torch.from_numpy(g["memories"])) < TOL[dtype]
    plan = enc._plan(*wave.shape)
    assert torch.equal(plan.view("frame_mask").cpu(), torch.from_numpy(g["frame_mask"]))
    assert plan.w2v_len64.tolist() == g["w2v_len"].tolist()
    assert rel_l2(plan.view("conv_feats").cpu()[:, ::37, ::11], torch.from_numpy(g["conv_feats_s"])) < TOL[dtype]
    assert rel_l2(plan.view("w2v_out").cpu()[:, ::11, ::37], torch.from_numpy(g["w2v_out_s"])) < TOL[dtype]
    assert rel_l2(plan.view("h_enc").cpu()[:, ::3, ::17], torch.from_numpy(g["h_enc_s"])) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_tiny64(dtype):
    g, wave, lens = _golden("tiny64")
    out = encoder(64, dtype)(wave.cuda(), lens.cuda())
    assert out.encoder_out.shape == (64, 3, 512) and out.encoder_padding_mask.shape == (3, 64)
    assert rel_l2(out.encoder_out.cpu(), torch.from_numpy(g["memories"])) < TOL[dtype]


def test_cuda_graph_replay_is_bit_identical_and_reusable():
    g, wave, lens = _golden("tiny")
    a = encoder(16, torch.float32, use_graph=False)(wave.cuda(), lens.cuda()).encoder_out
    eg = encoder(16, torch.float32, use_graph=True)
    b = eg(wave.cuda(), lens.cuda()).encoder_out
    assert torch.equal(a, b)
    wave2, lens2 = synth.make_waveforms([16000, 3000, 9999], seed=21)      # same shape, new lengths: replay
    c = eg(wave2.cuda(), lens2.cuda()).encoder_out
    with torch.no_grad():
        ref, _ = O.encoder_forward(synth.make_state_dict(seed=0), wave2, lens2)
    assert rel_l2(c.cpu(), ref) < 1e-5
    assert torch.equal(eg(wave.cuda(), lens.cuda()).encoder_out, a)          # and back again


def test_batch_composition_moves_with_the_reference():
    """the same utterance alone vs inside a padded batch gives different outputs in the
    reference (GroupNorm over padded time, ceil-style masks, unmasked memory attention); ours must follow."""
    sd = synth.make_state_dict(seed=0)
    wave, lens = synth.make_waveforms([24000, 9000], seed=5)
    enc = encoder(16, torch.float32)
    both = enc(wave.cuda(), lens.cuda()).encoder_out.cpu()
    alone = enc(wave[1:, :9000].contiguous().cuda(), lens[1:].cuda()).encoder_out.cpu()
    with torch.no_grad():
        r_both, _ = O.encoder_forward(sd, wave, lens)
        r_alone, _ = O.encoder_forward(sd, wave[1:, :9000].contiguous(), lens[1:])
    assert rel_l2(both, r_both) < 1e-5 and rel_l2(alone, r_alone) < 1e-5
    assert rel_l2(both[:, 1:], alone) > 1e-3            # they really do differ


def test_reorder_and_no_interlingua():
    g, wave, lens = _golden("tiny")
    enc = encoder(16, torch.float32)
    out = enc(wave.cuda(), lens.cuda())
    order = torch.tensor([2, 2, 0, 1], device="cuda")
    r = enc.reorder_encoder_out(out, order)
    assert torch.equal(r.encoder_out, out.encoder_out[:, order]) and r.encoder_padding_mask.shape == (4, 16)
    enc.no_interlingua = True
    try:
        h = enc(wave.cuda(), lens.cuda()).encoder_out
    finally:
        enc.no_interlingua = False
    assert rel_l2(h.cpu(), torch.from_numpy(g["h_enc"]).transpose(0, 1)) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_forward_many_on_stream_lanes_equals_forward(dtype):
    """Independent batches on concurrent CUDA-stream lanes: bit-identical to one-at-a-time forward()."""
    enc = encoder(16, dtype, use_graph=True)
    # (all batches wider than 64 wav2vec2 frames and narrower than 64 subsampled frames: cst_attention picks its kernel by the
    # launch's largest query count, so a super-batch mixing both sides of that threshold rounds differently from forward())
    shapes = [[26000, 22345, 8000], [29000, 7000], [24000], [36000, 3000, 9999], [29000, 8999], [32000, 11000, 10000, 500]]
    batches = []
    for i, lens in enumerate(shapes * 2):
        w, l = synth.make_waveforms(lens, seed=100 + i)
        batches.append((w.cuda(), l.cuda()))
    ref = [enc(w, l).encoder_out.clone() for w, l in batches]
    for lanes in (2, 3):
        got = enc.forward_many(batches, n_lanes=lanes)
        torch.cuda.synchronize()
        for a, b in zip(ref, got):
            assert torch.equal(a, b.encoder_out)
            assert b.encoder_padding_mask.shape == (a.shape[1], 16)
    # pinned HOST inputs / outputs (the end-to-end form bench.py times): copied on the lanes' own streams, same bits
    host = [(w.cpu().pin_memory(), l.cpu().pin_memory()) for w, l in batches]
    out_host = [torch.empty_like(r, device="cpu").pin_memory() for r in ref]
    enc.forward_many(host, n_lanes=3, out=out_host)
    torch.cuda.synchronize()
    for oh, r in zip(out_host, ref):
        assert torch.equal(oh, r.cpu())
    one = enc(host[0][0], host[0][1])
    assert one.encoder_out.is_cuda and torch.equal(one.encoder_out, ref[0])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_super_batch_small_groups_equal_forward(dtype):
    """Several reference batches of different padded widths in one row space (one launch sequence / one graph):
    bit-identical to forward() per batch, with and without lanes, for every grouping size."""
    enc = encoder(16, dtype, use_graph=True)
    shapes = [[26000, 22345, 8000], [29000, 7000], [24000], [36000, 3000, 9999], [29000, 8999], [32000, 11000, 10000, 500], [22000]]
    batches = []
    for i, lens in enumerate(shapes):
        w, l = synth.make_waveforms(lens, seed=200 + i)
        batches.append((w.cuda(), l.cuda()))
    ref = [enc(w, l).encoder_out.clone() for w, l in batches]
    for rows, lanes in ((1 << 30, 1), (150, 2), (0, 2)):          # one super-batch of 7 groups; several; none
        got = enc.forward_many(batches, n_lanes=lanes, super_rows=rows)
        torch.cuda.synchronize()
        for a, b in zip(ref, got):
            assert torch.equal(a, b.encoder_out), (rows, lanes, rel_l2(b.encoder_out, a))
    # groups on both sides of the attention dispatch threshold (<= 64 query rows -> the few-queries kernel): same
    # arithmetic, different rounding -- equal to forward() within the mode's tolerance, not bit for bit
    mixed = [synth.make_waveforms(lens, seed=250 + i) for i, lens in enumerate([[16000, 9000], [24000], [400]])]
    mixed = [(w.cuda(), l.cuda()) for w, l in mixed]
    ref = [enc(w, l).encoder_out.clone() for w, l in mixed]
    got = enc.forward_many(mixed, n_lanes=1, super_rows=1 << 30)
    for a, b in zip(ref, got):
        assert rel_l2(b.encoder_out, a) < (2e-5 if dtype == torch.float32 else 8e-3), rel_l2(b.encoder_out, a)


def test_super_batch_c3_shapes_equal_forward_bf16():
    """Real c3 shapes (2e6-sample token budget): four reference batches -> one super-batch of ~25k frame rows, where the
    GEMMs switch to the CTA-pair kernel (M >= 16384).  Memories must be those of forward() per batch."""
    enc = encoder(16, torch.bfloat16, use_graph=True)
    shapes = [[262960 - 777 * i for i in range(7)], [255000 - 500 * i for i in range(7)], [250000 - 300 * i for i in range(8)],
              [247000 - 311 * i for i in range(8)]]
    batches = []
    for i, lens in enumerate(shapes):
        w, l = synth.make_waveforms(lens, seed=300 + i)
        batches.append((w.cuda(), l.cuda()))
    ref = [enc(w, l).encoder_out.clone() for w, l in batches]
    got = enc.forward_many(batches, n_lanes=1)
    torch.cuda.synchronize()
    assert len(enc.plan_super_batches([tuple(w.shape) for w, _ in batches])) == 1
    for a, b in zip(ref, got):
        assert torch.equal(a, b.encoder_out), rel_l2(b.encoder_out, a)


def test_fast_fp32_mode_on_tensor_cores(monkeypatch):
    """CST_F32_TC=1: fp32 activations, GEMMs as 3-term fp16-split products on tcgen05 (fp32 accumulation in tensor memory).
    4x faster than the FFMA mode; the accumulation error of the tensor pipe (linear in K) puts it at ~3e-5, outside the 1e-5
    bar of the default fp32 mode, far inside the bf16 mode's 1e-2."""
    from chimera_st_b200.encoder import build_encoder_from_state_dict
    monkeypatch.setenv("CST_F32_TC", "1")
    g, wave, lens = _golden("c1mix")
    sd = synth.make_state_dict(seed=0, interlingua_length=16)
    enc = build_encoder_from_state_dict(sd, dtype=torch.float32, device="cuda", use_graph=False)
    out = enc(wave.cuda(), lens.cuda()).encoder_out.cpu()
    assert enc._plan(*wave.shape).f32_tc
    err = rel_l2(out, torch.from_numpy(g["memories"]))
    assert 1e-7 < err < 1e-4, err


@pytest.mark.parametrize("lever", ["CST_MEM_FUSED=1", "CST_LN_FUSE=0", "CST_LN_FUSE=2"])
def test_epilogue_levers_keep_parity_bf16(monkeypatch, lever):
    """The non-default forms of the LayerNorm / memory-stage plumbing (A/B levers) give the same memories within the
    16-bit tolerance: LN-fused weight-streaming linears in the memory stage, separate LayerNorm passes, LayerNorm fully fused around
    the GEMMs (the variant instantiations of the tcgen05 GEMM)."""
    from chimera_st_b200.encoder import build_encoder_from_state_dict
    k, v = lever.split("=")
    monkeypatch.setenv(k, v)
    g, wave, lens = _golden("c1mix")
    sd = synth.make_state_dict(seed=0, interlingua_length=16)
    enc = build_encoder_from_state_dict(sd, dtype=torch.bfloat16, device="cuda", use_graph=False)
    out = enc(wave.cuda(), lens.cuda()).encoder_out.float().cpu()
    assert rel_l2(out, torch.from_numpy(g["memories"])) < 1e-2


def test_modal_embedding_and_non_shared_encoder_layers_follow_the_reference_forward():
    """`modal_embedding` (row 0 for audio, row 1 for text, added to the memory queries; w2v2_transformer_interlingua.py:272-282) and
    `non_shared_encoder_layers` (audio runs audio_exclusive_layers[0..n) in place of transformer_layers[0..n), text keeps the shared
    ones; :239-249).  Checked against the oracle on state dicts rewritten to the equivalent plain model."""
    from chimera_st_b200.encoder import build_encoder_from_state_dict
    sd = synth.make_state_dict(seed=3, interlingua_length=16, text_vocab=50, modal_embedding=True, non_shared_encoder_layers=2)
    assert "modal_embedding.weight" in sd and "audio_exclusive_layers.1.fc2.bias" in sd
    enc = build_encoder_from_state_dict(sd, dtype=torch.float32, device="cuda", use_graph=False)
    assert set(enc.state_dict().keys()) == set(sd.keys())
    plain = {k: v for k, v in sd.items() if not k.startswith(("modal_embedding.", "audio_exclusive_layers."))}
    audio_eq, text_eq = dict(plain), dict(plain)
    audio_eq["interlingua_embedding.weight"] = sd["interlingua_embedding.weight"] + sd["modal_embedding.weight"][0]
    text_eq["interlingua_embedding.weight"] = sd["interlingua_embedding.weight"] + sd["modal_embedding.weight"][1]
    for k, v in sd.items():
        if k.startswith("audio_exclusive_layers."):
            audio_eq["transformer_layers." + k[len("audio_exclusive_layers."):]] = v
    wave, lens = synth.make_waveforms([12000, 7000], seed=4)
    tok = torch.randint(4, 50, (2, 9))
    tl = torch.tensor([9, 6])
    with torch.no_grad():
        ref_a, _ = O.encoder_forward(audio_eq, wave, lens)
        ref_t, _ = O.encoder_forward_text(text_eq, tok, tl)
    got_a = enc(wave.cuda(), lens.cuda()).encoder_out.cpu()
    got_t = enc(tok.cuda(), tl.cuda()).encoder_out.cpu()
    assert rel_l2(got_a, ref_a) < 1e-5 and rel_l2(got_t, ref_t) < 1e-5
    assert rel_l2(got_a, O.encoder_forward(plain, wave, lens)[0]) > 1e-3      # the options really change the result


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_base_encoder_mode_matches_the_reference_base_class(dtype):
    """enc.base_encoder = True: S2T_W2V2_TransformerEncoder.forward (w2v2_transformer.py:338-386) -- sinusoidal positions after
    the subsampler, LayerNorm-ed states as encoder_out, the real padding mask or None -- against goldens of the UNMODIFIED
    reference base-class forward (oracle/gen_golden_base.py)."""
    from chimera_st_b200.encoder import build_encoder_from_state_dict
    g = np.load(os.path.join(GOLDEN, "base_encoder.npz"))
    enc = build_encoder_from_state_dict(synth.make_state_dict(seed=0, interlingua_length=16), dtype=dtype, device="cuda", use_graph=False)
    enc.base_encoder = True
    for name in ("tiny", "full"):
        wave, lens = synth.make_waveforms(g[name + "_lens"].tolist(), seed=int(g[name + "_seed"]))
        out = enc(wave.cuda(), lens.cuda())
        ref = torch.from_numpy(g[name + "_encoder_out"])
        assert out.encoder_out.shape == ref.shape
        assert rel_l2(out.encoder_out.float().cpu(), ref) < TOL[dtype]
        if bool(g[name + "_has_mask"]):
            assert torch.equal(out.encoder_padding_mask.cpu(), torch.from_numpy(g[name + "_padding_mask"]))
        else:
            assert out.encoder_padding_mask is None
    enc.base_encoder = False
    mem = enc(wave.cuda(), lens.cuda())                      # back to the memory encoder: separate plan, [M, B, 512]
    assert mem.encoder_out.shape == (16, len(g["full_lens"]), 512)


def test_single_utterance_output_does_not_alias_the_arena():
    """B == 1: [1,M,512].transpose(0,1) is 'contiguous' to torch, so the result must be cloned explicitly."""
    enc = encoder(16, torch.float32, use_graph=True)
    w, l = synth.make_waveforms([24000], seed=102)
    a = enc(w.cuda(), l.cuda()).encoder_out
    keep = a.clone()
    w2, l2 = synth.make_waveforms([24000], seed=103)
    enc(w2.cuda(), l2.cuda())
    assert torch.equal(a, keep)


def test_16bit_mode_margin_and_pure_bf16_option():
    """Default 16-bit mode (bf16 + fp16 conv feature extractor) keeps a clear margin under the 1e-2 bar; the pure
    bf16 variant (conv_fp16=False) is still inside it."""
    from chimera_st_b200.encoder import build_encoder_from_state_dict
    g, wave, lens = _golden("tiny")
    sd = synth.make_state_dict(seed=0, interlingua_length=16)
    ref = torch.from_numpy(g["memories"])
    mixed = encoder(16, torch.bfloat16)(wave.cuda(), lens.cuda()).encoder_out.cpu()
    pure = build_encoder_from_state_dict(sd, dtype=torch.bfloat16, device="cuda", use_graph=False, conv_fp16=False)
    p = pure(wave.cuda(), lens.cuda()).encoder_out.cpu()
    assert rel_l2(mixed, ref) < 8e-3, rel_l2(mixed, ref)
    assert rel_l2(p, ref) < 1e-2, rel_l2(p, ref)


@pytest.mark.parametrize("B,secs,M", [(32, 15, 16), (64, 20, 64)])
def test_full_size_batches_row_independence_and_oracle_rows(B, secs, M):
    """BASELINE.json full sizes (C2: 32 x 15 s; C4: Chimera-64, 64 x 20 s, bf16).  The CPU oracle is too slow for
    the whole batch, so use a size-independent property: with equal padded length L every utterance's memories
    depend only on its own samples, hence (a) rows of the big batch equal the same utterance encoded alone
    (bit-identical: same kernels, row-wise arithmetic) and (b) two sampled rows match the oracle run on B=1."""
    L = secs * 16000
    lens = [L] + [int(L * (0.55 + 0.45 * ((7 * i) % 11) / 11.0)) for i in range(1, B)]
    wave, tl = synth.make_waveforms(lens, seed=300 + B)
    enc = encoder(M, torch.bfloat16, use_graph=True)
    out = enc(wave.cuda(), tl.cuda()).encoder_out
    assert out.shape == (M, B, 512) and bool(torch.isfinite(out).all())
    sd = synth.make_state_dict(seed=0, interlingua_length=M)
    for b in (0, B - 1):
        w1 = wave[b:b + 1].contiguous()
        l1 = torch.tensor([L])                      # same padded width: the row keeps its zero tail as signal
        alone = enc(w1.cuda(), l1.cuda()).encoder_out
        # frame masks differ (len == L alone vs lens[b] in the batch), so only row 0 (full length) is bit-comparable
        if lens[b] == L:
            assert torch.equal(alone[:, 0], out[:, b])
    w0, l0 = wave[0:1].contiguous(), torch.tensor([L])
    with torch.no_grad():
        ref, _ = O.encoder_forward(sd, w0, l0)
    assert rel_l2(out[:, 0:1].cpu(), ref) < 1e-2


def _cpu_reference(M):
    """The UNMODIFIED reference encoder when its tree or the shipped bundle (oracle/_ref/src) is there, else the oracle."""
    from oracle import make_overlay
    sd = synth.make_state_dict(seed=0, interlingua_length=M)
    if make_overlay.available():
        from oracle.ref_model import build_reference_encoder
        enc, _ = build_reference_encoder(M)
        enc.load_state_dict(sd, strict=True)
        return (lambda w, l: enc(w, l).encoder_out), "reference"
    return (lambda w, l: O.encoder_forward(sd, w, l)[0]), "oracle"


@pytest.mark.parametrize("case", ["c3_ragged_30s", "c3_ragged_short", "c2_full"])
def test_every_row_of_real_batch_shapes_bf16(case):
    """The benched configurations, EVERY row (incl. partially padded rows at T' = 1499 / 749): a c3 batch of 8 ragged
    utterances up to L = 480 000, a short-utterance c3 bucket, and the whole C2 batch (32 x 15 s, ragged) -- bf16 memories
    within 1e-2 rel-L2 per utterance of the fp32 reference on the identical padded batch."""
    torch.set_num_threads(os.cpu_count() or 8)
    if case == "c3_ragged_30s":
        lens = [480000, 479000, 471234, 455000, 430001, 401000, 377777, 350000]
    elif case == "c3_ragged_short":
        lens = [60000 - 1111 * i for i in range(24)]
    else:
        L = 240000
        lens = [L] + [int(L * (0.55 + 0.45 * ((7 * i) % 11) / 11.0)) for i in range(1, 32)]
    wave, tl = synth.make_waveforms(lens, seed=77)
    ref_fn, kind = _cpu_reference(16)
    with torch.no_grad():
        ref = ref_fn(wave, tl).double()                                  # [M, B, 512]
    enc = encoder(16, torch.bfloat16, use_graph=True)
    out = enc(wave.cuda(), tl.cuda()).encoder_out.cpu().double()
    per_row = (out - ref).pow(2).sum((0, 2)).sqrt() / ref.pow(2).sum((0, 2)).sqrt()
    assert float(per_row.max()) < 1e-2, (kind, per_row.tolist())
    assert rel_l2(out, ref) < 1e-2


@pytest.mark.parametrize("lens", [[400], [401, 400], [719, 500, 400], [960, 1], [3200, 3199, 17], [480000], [33000, 32000, 20000, 9000, 400]])
def test_edge_shapes_fp32(lens):
    """Ragged / extreme inputs the reference accepts: the shortest waveform the conv stack admits (L=400 -> T'=1),
    one-sample utterances inside a longer batch, the 30 s maximum of the workload, very uneven batches."""
    sd = synth.make_state_dict(seed=0)
    wave, tl = synth.make_waveforms(lens, seed=len(lens) * 13 + lens[0] % 97)
    st = {}
    with torch.no_grad():
        ref, _ = O.encoder_forward(sd, wave, tl, stages=st)
    enc = encoder(16, torch.float32)
    feat, fmask, flen = enc._get_w2v_feature(wave.cuda(), tl.cuda())
    assert torch.equal(fmask.cpu(), st["frame_mask"]) and torch.equal(flen.cpu(), st["w2v_len"])
    out = enc(wave.cuda(), tl.cuda())
    assert out.encoder_out.shape == ref.shape
    assert rel_l2(out.encoder_out.cpu(), ref) < 1e-5, rel_l2(out.encoder_out.cpu(), ref)
    assert rel_l2(feat.cpu(), st["w2v_out"]) < 1e-5


def test_rejects_bad_inputs():
    enc = encoder(16, torch.float32)
    with pytest.raises(ValueError):
        enc(torch.zeros(2, 3, 4000).cuda(), torch.tensor([4000, 4000]).cuda())
    with pytest.raises(ValueError):
        enc(torch.zeros(1, 100).cuda(), torch.tensor([100]).cuda())          # shorter than the conv stack's receptive field
    with pytest.raises(RuntimeError, match="text_embed_tokens"):          # text input needs the text embedding in the checkpoint
        enc(torch.zeros(1, 10, dtype=torch.long).cuda(), torch.tensor([10]).cuda())
    with pytest.raises(NotImplementedError):
        enc._get_w2v_feature(torch.zeros(1, 10, dtype=torch.long).cuda(), torch.tensor([10]).cuda())
