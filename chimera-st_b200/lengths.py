"""Integer side of the path: conv output lengths, the frame-level padding rule and the
subsampler lengths, in closed form (bit-exact against the reference's tensor code).

Reference behaviour restated:
  * conv stack lengths: T_i = (T_{i-1} - k)//s + 1, no padding (wav2vec2.py:707).
  * frame mask (wav2vec2.py:543-548 fed by lengths_to_padding_mask, data_utils.py:491-495):
    the [B,L] sample mask is trimmed by L % T', viewed [B,T',r] with r = L//T' and a frame is
    padded iff ALL its r samples are padded  =>  valid_b = min(T', ceil(len_b / r)).
    (ceil-style: 64000 samples of an 80000-wide batch give 200 frames although the conv
    stack itself would give 199 -- SURVEY.md §8(a) row a4.)
  * subsampler (s2t_transformer.py:63-67): l <- floor((l-1)/2 + 1) twice == (l+1)//2 for l>=0.
The CUDA kernel `cst_frame_lengths` evaluates the same formulas on the device so that the
forward pass needs no host sync (the reference does two `.item()` syncs here).
"""
from .synth import CONV_LAYERS


def conv_out_lengths(L):
    out = []
    for _, k, s in CONV_LAYERS:
        L = (L - k) // s + 1
        out.append(L)
    return out


def frame_valid_counts(src_lengths, L=None):
    """Valid (non-padded) wav2vec2 frames per utterance. src_lengths: iterable of ints."""
    lens = [int(x) for x in src_lengths]
    L = max(lens) if L is None else int(L)
    Tp = conv_out_lengths(L)[-1]
    if Tp <= 0:
        raise ValueError("input too short for the conv stack: L=%d" % L)
    r = L // Tp
    return [min(Tp, -(-n // r)) for n in lens]


def subsampler_len(l, n_layers=2):
    for _ in range(n_layers):
        l = (l + 1) // 2
    return l


def subsampler_out_frames(T, n_layers=2):
    """Conv1d(k=5,s=2,p=2) output length: (T + 4 - 5)//2 + 1 == (T+1)//2 for T>=1."""
    for _ in range(n_layers):
        T = (T + 4 - 5) // 2 + 1
    return T
