"""Host input pipeline: RIFF/WAVE PCM-16 reader against Python's own `wave` module and the
reference's float32 normalisation rule (soundfile: sample / 32768, fairseq/data/audio/audio_utils.py:33-55); the
16-bit-PCM-on-the-wire form of the encoder input (int16 src_tokens, scaled by 2^-15 on the device) on the ABI emulator
(CPU) and on the GPU."""
import io
import os
import struct
import wave as pywave

import numpy as np
import pytest
import torch

import chimera_st_b200  # noqa: F401
from chimera_st_b200 import audio_io, synth, weights, batching
from chimera_st_b200.plan import EncoderPlan
from conftest import rel_l2
from emu import EmuLib


def _wav_bytes(x, sr=16000, extra_chunk=False):
    bio = io.BytesIO()
    with pywave.open(bio, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes(np.asarray(x, dtype="<i2").tobytes())
    b = bio.getvalue()
    if extra_chunk:            # a LIST chunk (odd size -> pad byte) between fmt and data, as many encoders write
        lst = b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\x00"
        b = b[:36] + lst + b[36:]
        b = b[:4] + struct.pack("<I", len(b) - 8) + b[8:]
    return b


def test_reader_matches_python_wave_and_reference_normalisation(tmp_path):
    rng = np.random.RandomState(0)
    x = rng.randint(-32768, 32768, size=12345).astype(np.int16)
    for extra in (False, True):
        p = tmp_path / ("a%d.wav" % extra)
        p.write_bytes(_wav_bytes(x, extra_chunk=extra))
        got, sr = audio_io.read_pcm16(str(p))
        assert sr == 16000 and got.dtype == np.int16 and np.array_equal(got, x)
        f, sr = audio_io.get_waveform(str(p))
        assert f.dtype == np.float32 and np.array_equal(f, x.astype(np.float32) / np.float32(32768.0))   # libsndfile's rule
        f2, _ = audio_io.get_waveform(str(p), normalization=False)
        assert np.array_equal(f2, x.astype(np.float32))
    w, _ = audio_io.get_waveform_chi(str(p), 100, 50)
    assert np.array_equal(w, x[100:150].astype(np.float32) / np.float32(32768.0))
    # window forms of get_features_or_waveform: "<path>:<frame offset>:<frames>" and decimation by ori_sr // 16000
    assert np.array_equal(audio_io.get_features_or_waveform("%s:7:9" % p, pcm16=True), x[7:16])
    p48 = tmp_path / "b.wav"
    p48.write_bytes(_wav_bytes(x, sr=48000))
    assert np.array_equal(audio_io.get_features_or_waveform(str(p48), pcm16=True), x[::3])
    # a WAV stored uncompressed inside another file: "<zip path>:<byte offset>:<byte length>"
    blob = _wav_bytes(x[:777])
    pz = tmp_path / "c.zip"
    pz.write_bytes(b"PK" + b"\x00" * 30 + blob + b"tail")
    assert np.array_equal(audio_io.get_features_or_waveform("%s:32:%d" % (pz, len(blob)), pcm16=True), x[:777])
    with pytest.raises(ValueError):
        audio_io.get_waveform(str(tmp_path / "x.mp3"))
    with pytest.raises(FileNotFoundError):
        audio_io.get_features_or_waveform(str(tmp_path / "missing.wav"))
    pf = tmp_path / "d.flac"
    pf.write_bytes(b"fLaC" + b"\x00" * 40)
    with pytest.raises(NotImplementedError):
        audio_io.get_waveform(str(pf))
    bio = io.BytesIO()
    audio_io.write_pcm16(bio, x[:100])
    assert np.array_equal(audio_io.read_pcm16(bio.getvalue())[0], x[:100])


def _pcm_batch(lens, seed):
    g = torch.Generator().manual_seed(seed)
    waves = [(torch.randn(n, generator=g) * 3000).clamp(-32768, 32767).to(torch.int16) for n in lens]
    return batching.collate_waveforms(waves)


def test_collate_keeps_pcm16():
    ids, x, n = _pcm_batch([900, 1200, 400], 3)
    assert x.dtype == torch.int16 and x.shape == (3, 1200) and n.tolist() == [1200, 900, 400]
    assert not bool(x[2, 400:].any())


def test_int16_wire_equals_float_input_on_the_emulator():
    sd = synth.make_state_dict(seed=0, interlingua_length=16)
    P = weights.prepare(sd, torch.device("cpu"), torch.float32)
    _, x16, lens = _pcm_batch([5000, 4100], 5)
    xf = x16.float() / 32768.0
    a = EncoderPlan(P, 2, 5000, 16, torch.float32, torch.device("cpu"), lib=EmuLib())
    a.load_inputs(x16, lens)
    a.run()
    assert "wave_i16_to_f32" in a.lib.calls
    b = EncoderPlan(P, 2, 5000, 16, torch.float32, torch.device("cpu"), lib=EmuLib())
    b.load_inputs(xf, lens)
    b.run()
    assert torch.equal(a.waves[0], xf) and torch.equal(a.memories(), b.memories())


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_int16_wire_equals_float_input_on_the_gpu(dtype):
    """Pinned int16 host batch -> H2D (half the bytes) -> cst_wave_i16_to_f32 -> same bits as the float32 batch; also through
    forward_many (super-batch lanes) and from odd lengths (unaligned tails)."""
    from chimera_st_b200.encoder import build_encoder_from_state_dict
    sd = synth.make_state_dict(seed=0, interlingua_length=16)
    enc = build_encoder_from_state_dict(sd, dtype=dtype, device="cuda", use_graph=True)
    batches = [_pcm_batch(l, 7 + i)[1:] for i, l in enumerate([[16001, 12345, 8003], [9001, 7000], [3333]])]
    for x16, lens in batches:
        ref = enc((x16.float() / 32768.0).cuda(), lens.cuda()).encoder_out.clone()
        got = enc(x16.pin_memory(), lens).encoder_out
        assert got.dtype == torch.float32 and torch.equal(got, ref)
        assert torch.equal(enc(x16.cuda(), lens.cuda()).encoder_out, ref)
    many = enc.forward_many([(x.pin_memory(), l) for x, l in batches], n_lanes=2, super_rows=0)
    for (x16, lens), o in zip(batches, many):
        assert torch.equal(o.encoder_out, enc((x16.float() / 32768.0).cuda(), lens.cuda()).encoder_out)
