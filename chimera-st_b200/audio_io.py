"""Host-side waveform input: the reference's `get_waveform` / `get_features_or_waveform`
(fairseq/data/audio/audio_utils.py:33-55, fairseq/data/audio/speech_to_text_dataset.py:165-204) without soundfile.

The reference reads 16-bit mono WAV (or FLAC) through libsndfile as float32 = sample / 32768 and moves fp32 batches to
the GPU.  Here a RIFF/WAVE PCM-16 reader returns the int16 samples themselves (`read_pcm16`), which the encoder accepts
directly (`torch.int16` src_tokens: half the PCIe bytes, scaled by 2^-15 on the device -- `cst_wave_i16_to_f32`);
`get_waveform` / `get_features_or_waveform` keep the reference's signatures and float32 results for callers that want
them.  FLAC needs a decoder this image does not have: it raises, it is never silently mis-read.
"""
import io
import os
import struct
import zipfile

import numpy as np

_PCM, _EXTENSIBLE = 1, 0xFFFE


def _parse_riff(buf):
    """-> (fmt dict, data offset, data bytes).  Walks the chunk list (LIST / fact / cue chunks are skipped)."""
    if len(buf) < 12 or buf[0:4] != b"RIFF" or buf[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(buf):
        cid, size = buf[pos:pos + 4], struct.unpack_from("<I", buf, pos + 4)[0]
        body = pos + 8
        if cid == b"fmt ":
            tag, ch, sr, _, align, bits = struct.unpack_from("<HHIIHH", buf, body)
            if tag == _EXTENSIBLE and size >= 26:
                tag = struct.unpack_from("<H", buf, body + 24)[0]      # first two bytes of the sub-format GUID
            fmt = {"tag": tag, "channels": ch, "rate": sr, "align": align, "bits": bits}
        elif cid == b"data":
            data = (body, min(size, len(buf) - body))
            break
        pos = body + size + (size & 1)
    if fmt is None or data is None:
        raise ValueError("WAVE file without fmt / data chunk")
    return fmt, data[0], data[1]


def read_pcm16(path_or_fp, start=0, frames=-1):
    """16-bit PCM WAV -> (int16 ndarray [n] for mono / [n, channels], sample rate).  `start` / `frames` select a window
    in frames (the arguments soundfile.read takes in get_waveform_chi, audio_utils.py:8-30)."""
    if isinstance(path_or_fp, (bytes, bytearray, memoryview)):
        buf = bytes(path_or_fp)
    elif hasattr(path_or_fp, "read"):
        buf = path_or_fp.read()
    else:
        with open(path_or_fp, "rb") as f:
            buf = f.read()
    if buf[0:4] == b"fLaC":
        raise NotImplementedError("FLAC decoding is not available in this build (the reference uses libsndfile)")
    fmt, off, nbytes = _parse_riff(buf)
    if fmt["tag"] != _PCM or fmt["bits"] != 16:
        raise NotImplementedError("only 16-bit PCM WAV is supported (format tag %d, %d bits)" % (fmt["tag"], fmt["bits"]))
    ch = fmt["channels"]
    total = nbytes // (2 * ch)
    start = max(0, min(int(start), total))
    n = total - start if frames is None or frames < 0 else max(0, min(int(frames), total - start))
    x = np.frombuffer(buf, dtype="<i2", count=n * ch, offset=off + start * 2 * ch)
    return (x if ch == 1 else x.reshape(n, ch)), fmt["rate"]


def _check_ext(path_or_fp):
    if isinstance(path_or_fp, str):
        ext = os.path.splitext(os.path.basename(path_or_fp))[1]
        if ext not in {".flac", ".wav"}:
            raise ValueError(f"Unsupported audio format: {ext}")


def get_waveform(path_or_fp, normalization=True):
    """audio_utils.py:33-55: (float32 waveform, sample rate); normalised to [-1, 1) unless normalization=False."""
    _check_ext(path_or_fp)
    x, sr = read_pcm16(path_or_fp)
    w = x.astype(np.float32)
    if normalization:
        w *= np.float32(1.0 / 32768.0)
    return w, sr


def get_waveform_chi(path_or_fp, offset, length, normalization=True):
    """audio_utils.py:8-30: the same for a window of `length` frames starting at frame `offset`."""
    _check_ext(path_or_fp)
    x, sr = read_pcm16(path_or_fp, offset, length)
    w = x.astype(np.float32)
    if normalization:
        w *= np.float32(1.0 / 32768.0)
    return w, sr


def get_features_or_waveform(path, need_waveform=True, sample_rate=16000, pcm16=False):
    """speech_to_text_dataset.py:165-204 for waveform inputs: "<wav path>", "<wav path>:<frame offset>:<frames>" or
    "<zip path>:<byte offset>:<byte length>" (a WAV stored uncompressed in a ZIP); decimates by ori_sr // sample_rate
    exactly as the reference does (`wave[::k]`, no filtering).  pcm16=True returns the int16 samples (wire format)."""
    if not need_waveform:
        raise NotImplementedError("filter-bank features are not an input of the wav2vec2 path")
    _path, *extra = path.split(":")
    if not os.path.exists(_path):
        raise FileNotFoundError(f"File not found: {_path}")
    if len(extra) == 0:
        _check_ext(_path)
        x, ori_sr = read_pcm16(_path)
    elif len(extra) == 2:
        a, b = int(extra[0]), int(extra[1])
        if _path.endswith(".zip"):
            with open(_path, "rb") as f:
                f.seek(a)
                x, ori_sr = read_pcm16(f.read(b))
        else:
            _check_ext(_path)
            x, ori_sr = read_pcm16(_path, a, b)
    else:
        raise ValueError(f"Invalid path: {path}")
    if ori_sr != sample_rate:
        x = x[::(ori_sr // sample_rate)]
    if pcm16:
        return np.ascontiguousarray(x)
    return x.astype(np.float32) * np.float32(1.0 / 32768.0)


def write_pcm16(path_or_fp, samples, sample_rate=16000):
    """Minimal 16-bit mono WAV writer (tests / tools)."""
    x = np.asarray(samples, dtype="<i2")
    hdr = b"RIFF" + struct.pack("<I", 36 + x.nbytes) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, sample_rate,
                                                                                      2 * sample_rate, 2, 16)
    blob = hdr + b"data" + struct.pack("<I", x.nbytes) + x.tobytes()
    if hasattr(path_or_fp, "write"):
        path_or_fp.write(blob)
    else:
        with open(path_or_fp, "wb") as f:
            f.write(blob)


__all__ = ["read_pcm16", "get_waveform", "get_waveform_chi", "get_features_or_waveform", "write_pcm16"]
_ = (io, zipfile)
