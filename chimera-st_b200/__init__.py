"""chimera-st_b200: B200-native (sm_100a) speech-encoding hot path of Chimera-ST.

waveform -> wav2vec2 conv feature extractor -> pos-conv -> 12 wav2vec2 layers ->
conv subsampler -> 6 shared layers -> M shared semantic memories, as hand-written CUDA
kernels behind a C-ABI library (`include/chimera_st_b200.h`) with a thin PyTorch host
mirror of the reference encoder module.  There is no CPU fallback: importing the
compute modules without the built CUDA library raises.
"""
__version__ = "0.1.0"
from . import synth, lengths  # noqa: F401  (pure-host helpers; no CUDA needed)
