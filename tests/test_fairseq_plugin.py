"""The --user-dir drop-in, exercised against the UNMODIFIED reference (the mounted /root/reference in the dev container,
or the bundle oracle/_ref/src that `oracle/make_ref_bundle.py` ships to the GPU box).
CPU part: registry rebinding, model construction through the reference's own build_model, a strict state-dict exchange
with a reference model (encoder AND decoder keys), generator routing.
GPU part (-m gpu): the plugin model built through the registry, `model.half().cuda()` as generate.py:131-138 does, the
reference's OWN SequenceGenerator(beam 5) driving the B200 encoder, and the B200 generator, on one sample."""
import argparse
import os
import subprocess
import sys

import pytest

from conftest import ROOT

from oracle import make_overlay

pytestmark = pytest.mark.skipif(not make_overlay.available(), reason="no reference tree (mount or oracle/_ref/src bundle)")

SCRIPT = r'''
import sys, argparse, torch
sys.path.insert(0, %(root)r)
from oracle import make_overlay
make_overlay.build(); make_overlay.activate()
import fairseq.models
from fairseq import utils
from fairseq.models import ARCH_MODEL_REGISTRY
from fairseq.data import Dictionary
ref_cls = ARCH_MODEL_REGISTRY["s2t_transformer_w2v2_interlingua_base"]
utils.import_user_module(argparse.Namespace(user_dir=%(plugin)r))
new_cls = ARCH_MODEL_REGISTRY["s2t_transformer_w2v2_interlingua_base"]
assert new_cls is not ref_cls and issubclass(new_cls, ref_cls), (new_cls, ref_cls)
d = Dictionary.load(make_overlay.ref_root() + "/chimera/resources/wmt14-en-de-spm/spm_unigram10000_wave_joint.txt")
class Task: source_dictionary = None; target_dictionary = d
args = argparse.Namespace(w2v2_model_path="unused", encoder_layers=6, encoder_embed_dim=512, interlingua_length=16,
    interlingua_layers=3, interlingua_debug_options=[], dropout=0.1, share_decoder_input_output_embed=True,
    max_source_positions=6000, max_target_positions=1024)
model = new_cls.build_model(args, Task())
enc = model.encoder
import chimera_st_b200
from chimera_st_b200.encoder import B200InterlinguaEncoder
from chimera_st_b200 import synth
from fairseq.models import FairseqEncoder
assert isinstance(enc, B200InterlinguaEncoder) and isinstance(enc, FairseqEncoder)
sd = synth.make_state_dict(seed=0, interlingua_length=16)
full = {"encoder." + k: v for k, v in sd.items()}
full.update({k: v for k, v in model.state_dict().items() if k.startswith("decoder.")})
model.load_state_dict(full, strict=True)                      # fairseq_model.py:94-112 path (upgrade_state_dict runs)
assert torch.equal(model.encoder.interlingua_embedding.weight, sd["interlingua_embedding.weight"])
assert model.encoder.max_positions() is None
model.half()                                                  # generate.py:131-138: must not break the fp32 masters
assert model.encoder.layer_norm.weight.dtype == torch.float32
# generator routing (fairseq_task.py:309-412 hook): beam > 8 / sampling -> the reference SequenceGenerator; plain beam
# search with beam <= 8 (greedy included) -> the B200 generator, which refuses to be built on a CPU-resident model
from fairseq.tasks.fairseq_task import FairseqTask
from fairseq.sequence_generator import SequenceGenerator
from chimera_st_b200._lib import CstError
class T(FairseqTask):
    target_dictionary = d
    source_dictionary = None
task = T(argparse.Namespace())
model.float()
g16 = task.build_generator([model], argparse.Namespace(beam=16, controlled_generator=False))
assert type(g16) is SequenceGenerator, type(g16)
try:
    task.build_generator([model], argparse.Namespace(beam=5, controlled_generator=False))
    raise SystemExit("beam-5 generator was built on CPU")
except CstError as e:
    assert "CUDA" in str(e)
gs = task.build_generator([model], argparse.Namespace(beam=1, sampling=True, sampling_topk=3, controlled_generator=False))
assert type(gs) is SequenceGenerator, type(gs)
try:
    task.build_generator([model], argparse.Namespace(beam=1, controlled_generator=False))
    raise SystemExit("greedy generator was built on CPU")
except CstError as e:
    assert "CUDA" in str(e)
import os
os.environ["CHIMERA_B200_GREEDY"] = "0"
assert type(task.build_generator([model], argparse.Namespace(beam=1, controlled_generator=False))) is SequenceGenerator
print("PLUGIN_OK", type(enc).__name__, len(full))
'''


def test_user_dir_plugin_rebinds_arch_and_loads_reference_state_dict():
    code = SCRIPT % {"root": ROOT, "plugin": os.path.join(ROOT, "chimera-st_b200", "fairseq_plugin")}
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert "PLUGIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


GEN_SCRIPT = r'''
import sys, argparse, tempfile, os, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
from oracle import make_overlay
make_overlay.build(); make_overlay.activate()
import fairseq.models
from fairseq.data import Dictionary
from fairseq.models.wav2vec import wav2vec2 as W
from fairseq.models.chimera.w2v2_transformer_interlingua import S2TTransformerInterlinguaModelW2V2
from fairseq.sequence_generator import SequenceGenerator
from oracle.ref_model import W2V_CONV_SPEC
import chimera_st_b200
from chimera_st_b200 import synth
from chimera_st_b200.decoder import B200GreedyGenerator
from emu import EmuLib
torch.set_num_threads(8)
d = Dictionary.load(make_overlay.ref_root() + "/chimera/resources/wmt14-en-de-spm/spm_unigram10000_wave_joint.txt")
class Task: source_dictionary = None; target_dictionary = d
w2v_args = argparse.Namespace(conv_feature_layers=W2V_CONV_SPEC, quantize_targets=True, final_dim=256, encoder_layerdrop=0.05,
                              dropout_input=0.1, dropout_features=0.1, feature_grad_mult=0.1)
W.base_architecture(w2v_args)
tmp = tempfile.NamedTemporaryFile(suffix=".pt", delete=False); tmp.close()
torch.save({"args": w2v_args, "model": W.Wav2Vec2Model.build_model(w2v_args, task=None).state_dict()}, tmp.name)
args = argparse.Namespace(w2v2_model_path=tmp.name, encoder_layers=6, encoder_embed_dim=512, interlingua_length=16,
    interlingua_layers=3, interlingua_debug_options=[], dropout=0.1, share_decoder_input_output_embed=True,
    max_source_positions=6000, max_target_positions=1024)
model = S2TTransformerInterlinguaModelW2V2.build_model(args, Task()); os.unlink(tmp.name)
sd = {"encoder." + k: v for k, v in synth.make_state_dict(seed=0, interlingua_length=16).items()}
sd.update(synth.make_decoder_state_dict(seed=1))
sd["encoder.text_embed_tokens.weight"] = torch.zeros(synth.VOCAB, 512)
model.load_state_dict(sd, strict=True); model.eval()
wave, lens = synth.make_waveforms([9000, 6000], seed=5)
sample = {"net_input": {"src_tokens": wave, "src_lengths": lens}}
for beam, lenpen in ((1, 1.0), (4, 0.6)):
    kw = dict(beam_size=beam, max_len_a=0, max_len_b=7, len_penalty=lenpen)
    ref = SequenceGenerator([model], d, **kw).generate([model], sample)
    new = B200GreedyGenerator([model], d, lib=EmuLib(), **kw).generate([model], sample)     # reference encoder (CPU) + B200 search on the ABI emulator
    assert len(ref) == len(new) == 2
    for r, n in zip(ref, new):
        assert len(n) == len(r) == beam and set(n[0]) >= {"tokens", "score", "attention", "alignment", "positional_scores"}
        for rh, nh in zip(r, n):
            assert rh["tokens"].tolist() == nh["tokens"].tolist(), (beam, rh["tokens"], nh["tokens"])
            assert abs(float(rh["score"]) - nh["score"]) < 1e-4
            assert (rh["positional_scores"] - nh["positional_scores"]).abs().max() < 1e-4
print("GEN_OK", [h[0]["tokens"].tolist() for h in new])
'''


def test_greedy_generator_matches_reference_sequence_generator_structure_and_scores():
    """B200GreedyGenerator (decode logic on the host emulator of the C ABI) against the UNMODIFIED reference
    SequenceGenerator(beam_size=1) on the same reference model and sample: tokens, normalised score, positional scores."""
    code = GEN_SCRIPT % {"root": ROOT}
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=env)
    assert "GEN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


GPU_SCRIPT = r'''
import sys, argparse, tempfile, os, torch, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, %(root)r)
from oracle import make_overlay
make_overlay.build(); make_overlay.activate()
import fairseq.models
from fairseq import utils
from fairseq.models import ARCH_MODEL_REGISTRY
from fairseq.data import Dictionary
from fairseq.sequence_generator import SequenceGenerator
from fairseq.tasks.fairseq_task import FairseqTask
utils.import_user_module(argparse.Namespace(user_dir=%(plugin)r))           # generate.py:75
import chimera_st_b200
from chimera_st_b200 import synth, _lib
from chimera_st_b200.encoder import B200InterlinguaEncoder
from chimera_st_b200.decoder import B200GreedyGenerator
d = Dictionary.load(make_overlay.ref_root() + "/chimera/resources/wmt14-en-de-spm/spm_unigram10000_wave_joint.txt")
class T(FairseqTask):
    target_dictionary = d
    source_dictionary = None
task = T(argparse.Namespace())
sd = {"encoder." + k: v for k, v in synth.make_state_dict(seed=0, interlingua_length=16).items()}
dsd = synth.make_decoder_state_dict(seed=1)
dsd["decoder.embed_tokens.weight"][2] *= 3.0          # EOS row: hypotheses end at scattered steps (as tests/golden/beam.npz)
dsd["decoder.output_projection.weight"] = dsd["decoder.embed_tokens.weight"]
sd.update(dsd)
wave, lens = synth.make_waveforms([16000, 12345, 8000], seed=7)
KW = dict(max_len_a=0, max_len_b=12)
res = {}
for half in (False, True):
    args = argparse.Namespace(arch="s2t_transformer_w2v2_interlingua_base", w2v2_model_path="unused", encoder_layers=6,
        encoder_embed_dim=512, interlingua_length=16, interlingua_layers=3, interlingua_debug_options=[], dropout=0.1,
        share_decoder_input_output_embed=True, max_source_positions=6000, max_target_positions=1024, fp16=half)
    model = ARCH_MODEL_REGISTRY[args.arch].build_model(args, task)            # models/__init__.py:55-58, through the registry
    assert isinstance(model.encoder, B200InterlinguaEncoder)
    model.load_state_dict(sd, strict=True)
    model.eval()
    if half:
        model.half()                                                          # generate.py:131-134
    model.cuda()
    model.prepare_for_inference_(argparse.Namespace(generation=argparse.Namespace(beam=5, print_alignment=False)))   # :138
    assert model.encoder.layer_norm.weight.dtype == torch.float32 and model.encoder.layer_norm.weight.is_cuda
    x = wave.cuda().half() if half else wave.cuda()                           # generate.py:196-199 (move_to_cuda, apply_half)
    sample = {"net_input": {"src_tokens": x, "src_lengths": lens.cuda()}, "id": torch.arange(3)}
    # the encoder through the reference's own call seam (fairseq_encoder.py:43-62), with the collater's stray `mask=` kwarg
    eo = model.encoder.forward_torchscript({"src_tokens": x, "src_lengths": lens.cuda(), "mask": None})
    assert tuple(eo.encoder_out.shape) == (16, 3, 512) and eo.encoder_out.dtype == x.dtype
    pm = eo.encoder_padding_mask                                              # interlingua:301-312
    assert pm.dtype == torch.bool and tuple(pm.shape) == (3, 16) and not bool(pm.any())
    assert eo.encoder_embedding is None and eo.encoder_states is None
    ro = model.encoder.reorder_encoder_out(eo, torch.tensor([2, 0], device="cuda"))
    assert torch.equal(ro.encoder_out, eo.encoder_out[:, [2, 0]])
    # (1) the reference's OWN SequenceGenerator (beam 5) over the B200 encoder + the reference decoder modules
    ref_gen = SequenceGenerator([model], d, beam_size=5, **KW)
    ref = ref_gen.generate([model], sample)
    # (2) the generator the plugin routes --beam 5 to
    gen = task.build_generator([model], argparse.Namespace(beam=5, max_len_a=0, max_len_b=12, controlled_generator=False))
    assert isinstance(gen, B200GreedyGenerator), type(gen)
    new = gen.generate([model], sample)
    launches = model.encoder.last_launches
    assert launches > 100
    res[half] = (ref, new)
    for b, (r, n) in enumerate(zip(ref, new)):
        assert len(r) == len(n) == 5
        if not half:            # fp32: every hypothesis, in order, with its scores
            for rh, nh in zip(r, n):
                assert rh["tokens"].tolist() == nh["tokens"].tolist(), (b, rh["tokens"], nh["tokens"])
                assert abs(float(rh["score"]) - nh["score"]) < 2e-4
                assert (rh["positional_scores"].float().cpu() - nh["positional_scores"]).abs().max() < 2e-4
        else:                   # fp16 reference decoder vs bf16 B200 decoder: the best hypothesis, scores within half precision
            assert r[0]["tokens"].tolist() == n[0]["tokens"].tolist(), (b, r[0]["tokens"], n[0]["tokens"])
            assert abs(float(r[0]["score"]) - n[0]["score"]) < 5e-2
# 16-bit best hypotheses == fp32 best hypotheses
for b in range(3):
    assert res[True][1][b][0]["tokens"].tolist() == res[False][1][b][0]["tokens"].tolist()
print("GPU_PLUGIN_OK", [h[0]["tokens"].tolist() for h in res[False][1]], os.path.basename(_lib.LIB_PATH))
'''


@pytest.mark.gpu
def test_plugin_runs_under_real_fairseq_on_the_gpu():
    """utils.import_user_module -> registry -> model.half().cuda() -> the reference's own SequenceGenerator(beam 5) AND the
    B200 generator on one sample: identical tokens (fp32: all 5 hypotheses with scores), EncoderOut contract of
    w2v2_transformer_interlingua.py:301-312."""
    code = GPU_SCRIPT % {"root": ROOT, "plugin": os.path.join(ROOT, "chimera-st_b200", "fairseq_plugin")}
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=env)
    assert "GPU_PLUGIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
