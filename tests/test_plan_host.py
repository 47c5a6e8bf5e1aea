"""Host logic without a GPU: the real EncoderPlan launch sequence, weight preparation and frame
geometry, driven against the host emulator of the C ABI (tests/emu.py) and checked against the
oracle + golden vectors; plus state-dict layout and library-export checks."""
import os
import re

import numpy as np
import pytest
import torch

import chimera_st_b200  # noqa: F401
from chimera_st_b200 import synth, weights, _lib
from chimera_st_b200.plan import EncoderPlan, Geometry
from chimera_st_b200.encoder import B200InterlinguaEncoder, build_encoder_from_state_dict
from oracle import chimera_oracle as O
from conftest import rel_l2, GOLDEN, ROOT
from emu import EmuLib


@pytest.fixture(scope="module")
def tiny_plan():
    g = np.load(os.path.join(GOLDEN, "tiny.npz"))
    sd = synth.make_state_dict(seed=0, interlingua_length=16)
    wave, lens = synth.make_waveforms(g["src_lengths"].tolist(), seed=int(g["wave_seed"]))
    P = weights.prepare(sd, torch.device("cpu"), torch.float32)
    plan = EncoderPlan(P, wave.shape[0], wave.shape[1], 16, torch.float32, torch.device("cpu"), lib=EmuLib())
    plan.load_inputs(wave, lens)
    n = plan.run()
    return g, plan, n


def test_plan_sequence_matches_reference_goldens(tiny_plan):
    g, plan, n = tiny_plan
    assert n == len(plan.lib.calls) + 3          # conv0_stats is two kernels; + 2 memsets of the padded subsampler operands
    assert rel_l2(plan.view("conv_feats"), torch.from_numpy(g["conv_feats"])) < 5e-6
    assert rel_l2(plan.view("w2v_out"), torch.from_numpy(g["w2v_out"])) < 5e-6
    assert rel_l2(plan.view("h_enc"), torch.from_numpy(g["h_enc"])) < 5e-6
    assert rel_l2(plan.memories(), torch.from_numpy(g["memories"])) < 5e-6
    assert torch.equal(plan.view("frame_mask"), torch.from_numpy(g["frame_mask"]))
    assert torch.equal(plan.w2v_len64, torch.from_numpy(g["w2v_len"]))
    assert plan.sub_valid.tolist() == g["sub_len"].tolist()


def test_plan_is_reusable_with_other_lengths(tiny_plan):
    """Same (B, L) plan, different per-utterance lengths: padding buffers must not leak state."""
    _, plan, _ = tiny_plan
    sd = synth.make_state_dict(seed=0, interlingua_length=16)
    wave, lens = synth.make_waveforms([16000, 5000, 15999], seed=11)
    plan.load_inputs(wave, lens)
    plan.run()
    with torch.no_grad():
        ref, _ = O.encoder_forward(sd, wave, lens)
    assert rel_l2(plan.memories(), ref) < 5e-6


def test_super_batch_equals_each_batch_alone_on_the_emulator():
    """Three reference batches of different padded widths in ONE row space (shared row-wise launches, segment-table
    attention, per-group segment-aware launches): every group's stages, masks, lengths and memories equal the group run
    alone, and the oracle."""
    sd = synth.make_state_dict(seed=0, interlingua_length=16)
    P = weights.prepare(sd, torch.device("cpu"), torch.float32)
    batches = [[9000, 7000], [5000], [3300, 3000, 400]]
    data = [synth.make_waveforms(b, seed=21 + i) for i, b in enumerate(batches)]
    sup = EncoderPlan(P, None, None, 16, torch.float32, torch.device("cpu"), lib=EmuLib(),
                      groups=[tuple(w.shape) for w, _ in data])
    for k, (w, l) in enumerate(data):
        sup.load_inputs(w, l, group=k)
    sup.run()
    assert "attention_segs" in sup.lib.calls and "attention" not in sup.lib.calls
    for k, (w, l) in enumerate(data):
        one = EncoderPlan(P, w.shape[0], w.shape[1], 16, torch.float32, torch.device("cpu"), lib=EmuLib())
        one.load_inputs(w, l)
        one.run()
        for name in ("conv_feats", "w2v_out", "h_enc"):
            assert rel_l2(sup.view(name, k), one.view(name)) < 2e-6, (k, name)
        assert torch.equal(sup.view("frame_mask", k), one.view("frame_mask"))
        assert torch.equal(sup.view("w2v_len64", k), one.w2v_len64)
        assert rel_l2(sup.memories(k), one.memories()) < 2e-6
        with torch.no_grad():
            ref, _ = O.encoder_forward(sd, w, l)
        assert rel_l2(sup.memories(k), ref) < 5e-6


@pytest.mark.parametrize("L", [400, 401, 719, 720, 16000, 80000, 240000, 480000])
def test_geometry_invariants(L):
    g = Geometry(3, L, 16)
    assert g.Tp == O.conv_out_lengths(L)[-1]
    for i in range(7):
        assert g.Ta[i] >= g.T[i]
        if i:
            assert g.Ta[i] * 2 == g.Ta[i - 1]
            # a valid output frame never reads beyond the valid frames of the level below
            k = synth.CONV_LAYERS[i][1]
            assert 2 * (g.T[i] - 1) + k - 1 <= g.T[i - 1] - 1
    assert g.T1 == (g.Tp + 4 - 5) // 2 + 1 and g.T2 == (g.T1 + 4 - 5) // 2 + 1
    assert g.Tin1 % 2 == 0 and g.Tin1 >= g.Tp + 4 and g.Tin2 % 2 == 0 and g.Tin2 >= g.T1 + 4
    assert 2 * (g.T1 - 1) + 4 <= g.Tin1 - 1 and 2 * (g.T2 - 1) + 4 <= g.Tin2 - 1


def test_state_dict_layout_and_strict_load():
    sd = synth.make_state_dict(seed=3, interlingua_length=16)
    enc = B200InterlinguaEncoder(16)
    assert list(enc.state_dict().keys()) == list(sd.keys()) or set(enc.state_dict()) == set(sd)
    for k, v in enc.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    enc.load_state_dict(sd, strict=True)
    enc2 = build_encoder_from_state_dict({"encoder." + k: v for k, v in sd.items()}, device="cpu")
    assert torch.equal(enc2.state_dict()["interlingua_embedding.weight"], sd["interlingua_embedding.weight"])
    assert enc2.max_positions() is None
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc2(torch.zeros(1, 16000), torch.tensor([16000]))           # product path refuses to run on CPU
    with pytest.raises(RuntimeError, match="no CPU fallback"):         # text branch: same rule
        enc2(torch.zeros(1, 10, dtype=torch.long), torch.tensor([10]))
    with pytest.raises(NotImplementedError):
        enc2._get_w2v_feature(torch.zeros(1, 10, dtype=torch.long), torch.tensor([10]))


def test_posconv_weight_folding_accepts_both_forms():
    sd = synth.make_state_dict(seed=0, interlingua_length=16)
    P1 = weights.prepare(sd, torch.device("cpu"), torch.float32)
    sd2 = dict(sd)
    pc = "wav2vec_model.encoder.pos_conv.0."
    sd2[pc + "weight"] = O.pos_conv_weight(sd)
    del sd2[pc + "weight_g"], sd2[pc + "weight_v"]
    P2 = weights.prepare(sd2, torch.device("cpu"), torch.float32)
    assert torch.allclose(P1["pos_w"], P2["pos_w"], rtol=0, atol=0)


def test_library_loads_and_exports_every_declared_symbol():
    """The C-ABI library must be built in-tree and export exactly what include/*.h declares."""
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "chimera_st_b200.h")).read()
    declared = set(re.findall(r"\b(cst_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.cst_abi_version() == 1
    assert lib.cst_last_error() is not None


def test_gemm_rejects_bad_arguments_without_touching_the_device():
    import ctypes as C
    lib = _lib.load()
    p = _lib.GemmParams()
    assert lib.cst_gemm(C.byref(p), None) != 0
    assert b"null pointer" in lib.cst_last_error()


def test_ctypes_signatures_agree_with_the_header_prototypes():
    """Every prototype in include/chimera_st_b200.h against the ctypes signature the host side binds (_lib._SIGS): same number of
    arguments, and per argument the same class (pointer / float / 64-bit integer / 32-bit integer) -- ABI drift between the header, the
    library and the binding shows up here, without a GPU."""
    import ctypes as C
    hdr = open(os.path.join(ROOT, "include", "chimera_st_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", "", hdr)
    protos = re.findall(r"\b(?:int|long long|const char\*|void)\s+(cst_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)
    assert len(protos) >= 40

    def klass_c(arg):
        arg = " ".join(arg.split())
        if "*" in arg:
            return "ptr"
        base = arg.rsplit(" ", 1)[0] if " " in arg else arg
        if base in ("float",):
            return "f32"
        if base in ("long long", "int64_t", "unsigned long long", "size_t"):
            return "i64"
        if base in ("int", "unsigned int", "unsigned", "int32_t", "uint32_t"):
            return "i32"
        raise AssertionError("unclassified header argument: %r" % arg)

    def klass_py(t):
        if t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or getattr(t, "_type_", None) is not None and issubclass(t, C._Pointer):
            return "ptr"
        if t is C.c_float:
            return "f32"
        if t in (C.c_longlong, C.c_ulonglong, C.c_int64, C.c_uint64, C.c_size_t):
            return "i64"
        if t in (C.c_int, C.c_uint, C.c_int32, C.c_uint32):
            return "i32"
        raise AssertionError("unclassified ctypes argument: %r" % (t,))
    checked = 0
    for name, args in protos:
        if name not in _lib._SIGS:
            continue
        _, argtypes = _lib._SIGS[name]
        cargs = [a for a in (x.strip() for x in args.split(",")) if a and a != "void"]
        assert len(cargs) == len(argtypes), (name, len(cargs), len(argtypes))
        got = [klass_py(t) for t in argtypes]
        want = [klass_c(a) for a in cargs]
        assert got == want, (name, [(i, w, g) for i, (w, g) in enumerate(zip(want, got)) if w != g])
        checked += 1
    assert checked == len(_lib._SIGS) == len(protos), (checked, len(_lib._SIGS), len(protos))
