"""fairseq `--user-dir` plugin: drops the B200 encoder under the reference's unmodified CLIs.

    fairseq-generate <data> --user-dir /path/to/repo/chimera-st_b200/fairseq_plugin --path ckpt.pt ...
    fairseq-interactive ...   (same flag)

`fairseq.utils.import_user_module` (fairseq/utils.py:431-459) imports this package before the model is
built (generate.py:75, interactive.py:115, train.py:55).  A checkpoint stores
`args.arch = "s2t_transformer_w2v2_interlingua_base"`; we rebind the registries
(fairseq/models/__init__.py:31-165) so that arch resolves to a model class whose `build_encoder`
(w2v2_transformer_interlingua.py:125-130) returns `B200InterlinguaEncoder` -- same parameter names, so
`model.load_state_dict(checkpoint["model"])` works unchanged.  The decoder module stays the reference's (it owns the
`decoder.*` parameters); with `--beam` <= 8 (fairseq's default is 5) and otherwise default search options
`task.build_generator` returns the B200 generator (chimera_st_b200/decoder.py: greedy for beam 1, beam search for 2..8),
which decodes from those parameters on the GPU -- any other search setting, or CHIMERA_B200_SEARCH=0 (alias CHIMERA_B200_GREEDY=0), keeps the
reference's SequenceGenerator.

Compute dtype: `--fp16` / `--bf16` / `--memory-efficient-*` select the bf16 tensor-core path, otherwise
fp32; override with CHIMERA_B200_DTYPE=fp32|bf16.  Requires a CUDA device (there is no CPU fallback).
"""
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(os.path.dirname(_HERE))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

import chimera_st_b200  # noqa: E402,F401
from chimera_st_b200.encoder import B200InterlinguaEncoder  # noqa: E402

from fairseq.models import MODEL_REGISTRY, ARCH_MODEL_REGISTRY, FairseqEncoder  # noqa: E402
from fairseq.models.chimera.w2v2_transformer_interlingua import S2TTransformerInterlinguaModelW2V2  # noqa: E402


def _compute_dtype(args):
    env = os.environ.get("CHIMERA_B200_DTYPE", "")
    if env:
        return {"fp32": torch.float32, "bf16": torch.bfloat16}[env]
    half = any(getattr(args, k, False) for k in ("fp16", "bf16", "memory_efficient_fp16", "memory_efficient_bf16"))
    return torch.bfloat16 if half else torch.float32


class B200FairseqEncoder(B200InterlinguaEncoder, FairseqEncoder):
    """B200InterlinguaEncoder with the FairseqEncoder mix-in (forward_torchscript, set_num_updates, ...)."""

    def __init__(self, args, src_dict=None, embed_tokens=None):
        dbg = list(getattr(args, "interlingua_debug_options", []) or [])
        if [o for o in dbg if o != "modal_embedding"]:
            raise NotImplementedError("interlingua_debug_options %s are not supported by the B200 path" % dbg)
        vocab = embed_tokens.num_embeddings if embed_tokens is not None else 0
        B200InterlinguaEncoder.__init__(self, interlingua_length=args.interlingua_length, dtype=_compute_dtype(args),
                                        use_graph=True, dead_heads=True, text_vocab=vocab, modal_embedding="modal_embedding" in dbg,
                                        non_shared_encoder_layers=int(getattr(args, "non_shared_encoder_layers", 0) or 0))
        self.dictionary = src_dict
        self.no_interlingua = getattr(args, "no_interlingua", False)

    # fairseq's model.half()/bfloat16() (generate.py:131-138) must not down-cast the fp32 master parameters:
    # the kernels pick their own operand precision (weights.prepare).  Device moves still apply.
    def _apply(self, fn, recurse=True):
        return super()._apply(lambda t: fn(t).to(t.dtype) if t.is_floating_point() else fn(t), recurse)


class B200S2TInterlinguaModel(S2TTransformerInterlinguaModelW2V2):
    @classmethod
    def build_encoder(cls, args, src_dict=None, encoder_embed_tokens=None):
        return B200FairseqEncoder(args, src_dict, encoder_embed_tokens)


def register():
    """Rebind every registry entry of the reference interlingua model to the B200-backed class."""
    n = 0
    for name, klass in list(MODEL_REGISTRY.items()):
        if klass is S2TTransformerInterlinguaModelW2V2:
            MODEL_REGISTRY[name] = B200S2TInterlinguaModel
            n += 1
    for arch, klass in list(ARCH_MODEL_REGISTRY.items()):
        if klass is S2TTransformerInterlinguaModelW2V2:
            ARCH_MODEL_REGISTRY[arch] = B200S2TInterlinguaModel
            n += 1
    return n


def _plain_greedy(args, seq_gen_cls):
    flags = ("score_reference", "sampling", "constraints", "print_alignment", "match_source_len", "unnormalized",
             "controlled_generator")                 # the Chimera fork's own generator switch (fairseq_task.py:392)
    return (seq_gen_cls is None and 1 <= getattr(args, "beam", 5) <= 8 and not any(getattr(args, k, False) for k in flags)
            and getattr(args, "diverse_beam_groups", -1) <= 0 and getattr(args, "diversity_rate", -1) <= -1      # > -1 selects DiverseSiblingsSearch (fairseq_task.py:349)
            and getattr(args, "no_repeat_ngram_size", 0) == 0 and getattr(args, "temperature", 1.0) == 1.0
            and getattr(args, "unkpen", 0) == 0 and getattr(args, "prefix_size", 0) == 0
            and getattr(args, "prefix_allowed_tokens_fn", None) is None)


def patch_build_generator():
    """`--beam` <= 8 with default search options -> B200GreedyGenerator (FairseqTask.build_generator,
    fairseq/tasks/fairseq_task.py:309-412, is what every task's override ends in)."""
    from fairseq.tasks.fairseq_task import FairseqTask
    from chimera_st_b200.decoder import B200GreedyGenerator
    if getattr(FairseqTask.build_generator, "_b200", False):
        return
    orig = FairseqTask.build_generator

    def build_generator(self, models, args, seq_gen_cls=None, extra_gen_cls_kwargs=None):
        on = os.environ.get("CHIMERA_B200_SEARCH", os.environ.get("CHIMERA_B200_GREEDY", "1")) != "0"
        if (on and _plain_greedy(args, seq_gen_cls) and len(models) == 1
                and isinstance(models[0], B200S2TInterlinguaModel)):
            extra = extra_gen_cls_kwargs or {}
            try:
                return B200GreedyGenerator(models, self.target_dictionary, beam_size=getattr(args, "beam", 5),
                                           max_len_a=getattr(args, "max_len_a", 0), max_len_b=getattr(args, "max_len_b", 200),
                                           min_len=getattr(args, "min_len", 1), len_penalty=getattr(args, "lenpen", 1),
                                           symbols_to_strip_from_output=extra.get("symbols_to_strip_from_output"))
            except NotImplementedError as e:       # a decoder variant / option the B200 decoder does not cover
                print("chimera_st_b200: keeping the reference SequenceGenerator (%s)" % e, file=sys.stderr)
        return orig(self, models, args, seq_gen_cls=seq_gen_cls, extra_gen_cls_kwargs=extra_gen_cls_kwargs)
    build_generator._b200 = True
    FairseqTask.build_generator = build_generator


register()
patch_build_generator()
