"""Host-side mirror of the reference encoder module for the audio branch.

`B200InterlinguaEncoder` keeps the interface of
`S2T_W2V2_TransformerInterlinguaEncoder` (fairseq/models/chimera/w2v2_transformer_interlingua.py:155-312):

    forward(src_tokens [B,L] float, src_lengths [B] long, **extra) -> EncoderOut
    _get_w2v_feature(src_tokens, src_lengths) -> (feature [B,T',768], padding_mask [B,T'] bool, lengths [B] long)
    reorder_encoder_out(encoder_out, new_order), max_positions() -> None, upgrade_state_dict_named(sd, name)

and the SAME parameter names / shapes (state-dict layout of SURVEY.md App. D, incl. the dead
pre-training heads and `embed_positions._float_tensor`), so a reference checkpoint loads with
`load_state_dict(strict=True)`.  The arithmetic is done only by the CUDA kernels behind the C ABI
(`include/chimera_st_b200.h`); without the built library or without a CUDA device `forward` raises.

Integer (int32 / int64) `src_tokens` take the text (MT) branch (embedding + sinusoidal positions -> the same shared layers
and memory stage).  `torch.int16` src_tokens are 16-bit PCM samples "on the wire" (SURVEY §8 f.3): copied as int16 and
scaled by 2^-15 on the device, bit-identical to the reference's float32 waveform read at half the PCIe bytes.
`modal_embedding=True` (the reference's `interlingua_debug_options`) and `non_shared_encoder_layers=n` register the reference's extra
parameters (`modal_embedding.weight`, `audio_exclusive_layers.*`) and follow its forward (:239-249, :272-282).
Not reproduced (out of scope): training-time dropout / LayerDrop in `forward` (the training step lives in train.py).
"""
import os
from collections import OrderedDict
from typing import List, NamedTuple, Optional

import torch
import torch.nn as nn

from . import weights as _weights
from .plan import EncoderPlan, TextPlan, Arena
from .synth import encoder_param_spec, ENC_DIM


class EncoderOut(NamedTuple):
    """Field-for-field copy of fairseq.models.fairseq_encoder.EncoderOut (fairseq_encoder.py:13-23)."""
    encoder_out: torch.Tensor                              # M x B x C
    encoder_padding_mask: Optional[torch.Tensor]           # B x M, all False (not None: interlingua:301-305)
    encoder_embedding: Optional[torch.Tensor]
    encoder_states: Optional[List[torch.Tensor]]
    src_tokens: Optional[torch.Tensor]
    src_lengths: Optional[torch.Tensor]


class _Node(nn.Module):
    """Anonymous container: only there to give parameters the reference's dotted names."""


def _register(root, dotted, tensor, buffer=False):
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if not hasattr(mod, p):
            mod.add_module(p, _Node())
        mod = getattr(mod, p)
    if buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


class B200InterlinguaEncoder(nn.Module):
    MAX_PLANS = 256         # cached (B, L) shapes (geometry + CUDA graph); activations overlay one shared arena

    def __init__(self, interlingua_length=16, dtype=torch.float32, use_graph=True, dead_heads=True,
                 text_vocab=0, encoder_out_dtype=None, conv_fp16=None, modal_embedding=False, non_shared_encoder_layers=0):
        nn.Module.__init__(self)          # explicit: the fairseq plugin mixes this class with FairseqEncoder
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("compute dtype must be float32 or bfloat16")
        self.interlingua_length = interlingua_length
        self.compute_dtype = dtype
        # 16-bit mode: fp16 operands inside the normalisation-free conv feature extractor (see weights.py); bf16 elsewhere
        if conv_fp16 is None:
            conv_fp16 = os.environ.get("CST_CONV_FP16", "1") != "0"
        self.conv_dtype = torch.float16 if (dtype == torch.bfloat16 and conv_fp16) else dtype
        self.use_graph = use_graph
        self.encoder_out_dtype = encoder_out_dtype
        self.no_interlingua = False
        # base_encoder = True: the arithmetic of the reference's NON-memory encoder (S2T_W2V2_TransformerEncoder.forward,
        # fairseq/models/chimera/w2v2_transformer.py:338-386) on the same parameters: sinusoidal positions after the subsampler,
        # encoder_out = LayerNorm-ed states [T2, B, 512], the real key-padding mask (None when nothing is padded)
        self.base_encoder = False
        for name, shape, _, _ in encoder_param_spec(interlingua_length, dead_heads, text_vocab, modal_embedding, non_shared_encoder_layers):
            _register(self, name, torch.zeros(shape), buffer=name.endswith("_float_tensor"))
        self._prepared = None
        self._plans = OrderedDict()
        self._arena = None
        self._lanes = None          # extra (stream, arena, plans) lanes for forward_many
        self.last_launches = 0
        self.register_load_state_dict_post_hook(lambda m, k: m.invalidate())

    # ---- reference-compatible surface ------------------------------------------------------------
    def max_positions(self):
        return None                                           # interlingua:204-205

    def upgrade_state_dict_named(self, state_dict, name):
        key = name + ".text_embed_tokens.weight"              # interlingua:198-202
        if key in state_dict and not hasattr(self, "text_embed_tokens"):
            state_dict.pop(key)
        return state_dict

    def reorder_encoder_out(self, encoder_out, new_order):
        """w2v2_transformer.py:388-429 (beam replication): index_select on dim 1 / dim 0."""
        eo = encoder_out.encoder_out.index_select(1, new_order)
        pm = encoder_out.encoder_padding_mask
        pm = pm if pm is None else pm.index_select(0, new_order)
        return EncoderOut(eo, pm, None, None, None, None)

    def invalidate(self):
        """Drop prepared weights and cached plans (call after mutating parameters in place)."""
        self._prepared = None
        self._plans.clear()
        self._arena = None
        self._lanes = None

    # ---- plumbing ------------------------------------------------------------------------------------
    def _device(self):
        dev = self.layer_norm.weight.device
        if dev.type != "cuda":
            raise RuntimeError("B200InterlinguaEncoder runs only on a CUDA device (no CPU fallback); call .cuda()")
        return dev

    def _plan(self, B, L, lane=None, text=False, groups=None):
        dev = self._device()
        if self._prepared is None:
            self._prepared = _weights.prepare(self.state_dict(), dev, self.compute_dtype, self.conv_dtype)
        if lane is None:
            if self._arena is None:
                self._arena = Arena(dev)
            plans, arena = self._plans, self._arena
        else:
            plans, arena = lane["plans"], lane["arena"]
        base = bool(self.base_encoder) and not text
        key = (B, L, text, base) if groups is None else tuple(groups) + (base,)
        plan = plans.get(key)
        if plan is None:
            while len(plans) >= self.MAX_PLANS:
                plans.popitem(last=False)
            gen = arena.generation
            if text:
                plan = TextPlan(self._prepared, B, L, self.interlingua_length, self.compute_dtype, dev, self.use_graph,
                                arena=arena)
            else:
                plan = EncoderPlan(self._prepared, B, L, self.interlingua_length, self.compute_dtype, dev, self.use_graph,
                                   arena=arena, conv_dtype=self.conv_dtype, groups=groups, audio_positions=base)
            if arena.generation != gen:                # arena grew: older plans (and their graphs) point at freed memory
                plans.clear()
            plans[key] = plan
        else:
            plans.move_to_end(key)
        return plan

    def _get_lanes(self, n):
        dev = self._device()
        if self._lanes is None or len(self._lanes) != n:
            self._lanes = [{"stream": torch.cuda.Stream(device=dev), "arena": Arena(dev), "plans": OrderedDict()}
                           for _ in range(n)]
        return self._lanes

    SUPER_ROWS = 49152      # forward_many: wav2vec2 frame rows per super-batch (measured sweet spot on c3: 24576 / 36864 / 49152 / 65536 / 98304 rows = 188.0 / 187.6 / 181.9-183.8 / 184.4 / 183.9 ms per step); 0 = one plan per batch
    SUPER_MAX_GROUPS = 16

    @staticmethod
    def frame_rows(B, L):
        """Rows a padded [B, L] batch occupies in the wav2vec2 stage (allocation grid of plan.Geometry)."""
        t0 = (L - 10) // 5 + 1
        return B * ((t0 + 63) // 64)

    def plan_super_batches(self, shapes, super_rows=None):
        """Consecutive batches [(B, L), ...] -> index lists, each closed once it holds >= super_rows frame rows."""
        super_rows = self.SUPER_ROWS if super_rows is None else super_rows
        out, cur, rows = [], [], 0
        for i, (B, L) in enumerate(shapes):
            cur.append(i)
            rows += self.frame_rows(B, L)
            if super_rows <= 0 or rows >= super_rows or len(cur) >= self.SUPER_MAX_GROUPS:
                out.append(cur)
                cur, rows = [], 0
        if cur:
            out.append(cur)
        return out

    @torch.no_grad()
    def forward_many(self, batches, n_lanes=2, out=None, super_rows=None):
        """Throughput API: encode a list of independent padded batches [(src_tokens, src_lengths), ...].
        Consecutive batches are packed into SUPER-BATCHES of >= `super_rows` wav2vec2 frame rows: each reference batch
        keeps its own padded width, GroupNorm extent and frame mask (results are those of calling forward() per batch),
        but all of them share one row space, so the row-wise GEMM / LayerNorm launches see >= 24k rows instead of the
        ~6k of one 2e6-sample batch, and one CUDA graph covers the whole super-batch.  Super-batches are issued
        round-robin on `n_lanes` CUDA streams, each lane with its own activation arena and graphs.
        Returns a list of EncoderOut; `out` (optional list of preallocated [M,B,512] tensors, e.g. pinned host
        buffers) receives the memories with non_blocking copies.  src_tokens / src_lengths may be (pinned) HOST
        tensors: they are copied straight into the lane's input buffers on the lane's stream, so one super-batch's H2D
        transfer overlaps the other lanes' kernels."""
        lanes = self._get_lanes(n_lanes)
        cur = torch.cuda.current_stream()
        for ln in lanes:
            ln["stream"].wait_stream(cur)
        for src_tokens, src_lengths in batches:
            self._check_inputs(src_tokens, src_lengths)
        results = [None] * len(batches)
        self.last_launches = 0
        supers = self.plan_super_batches([tuple(b[0].shape) for b in batches], super_rows)
        for si, idx in enumerate(supers):
            ln = lanes[si % n_lanes]
            with torch.cuda.stream(ln["stream"]):
                groups = [tuple(batches[i][0].shape) for i in idx]
                plan = (self._plan(groups[0][0], groups[0][1], ln) if len(groups) == 1
                        else self._plan(None, None, ln, groups=groups))
                for k, i in enumerate(idx):
                    src_tokens, src_lengths = batches[i]
                    plan.load_inputs(src_tokens if src_tokens.dtype in (torch.float32, torch.int16) else src_tokens.float(),
                                     src_lengths, group=k)
                self.last_launches += plan.run()
                for k, i in enumerate(idx):
                    src_tokens = batches[i][0]
                    o = plan.memories(k).to(self.encoder_out_dtype or self._out_dtype(src_tokens)).clone(memory_format=torch.contiguous_format)
                    if out is not None:
                        out[i].copy_(o, non_blocking=True)
                    pad = torch.zeros(src_tokens.shape[0], o.shape[0], dtype=torch.bool, device=o.device)
                    results[i] = EncoderOut(o, pad, None, None, None, None)
        for ln in lanes:
            cur.wait_stream(ln["stream"])
        return results

    @staticmethod
    def _out_dtype(src_tokens):
        """The reference returns the model dtype, which equals the waveform dtype; int16 PCM input -> float32."""
        return src_tokens.dtype if src_tokens.dtype.is_floating_point else torch.float32

    def _check_inputs(self, src_tokens, src_lengths, allow_text=False):
        if not src_tokens.dtype.is_floating_point and src_tokens.dtype != torch.int16 and not allow_text:
            raise NotImplementedError("integer (text) tokens are only accepted by forward()")
        if self.training:
            raise NotImplementedError("the module's forward is the eval() path; the training step (dropout, LayerDrop, backward) is "
                                      "chimera_st_b200.train.EncoderTrainStep")
        if src_tokens.dim() != 2 or src_lengths.shape != (src_tokens.shape[0],):
            raise ValueError("expected src_tokens [B,L], src_lengths [B]")
        # Precondition of the frame-mask rule (a4): the batch is padded exactly to its longest utterance, as the
        # reference's collater guarantees (speech_to_text_dataset.py:218).  The reference derives the mask width from
        # max(src_lengths) (w2v2_transformer.py:327) and breaks on an over-padded batch; we use L = src_tokens.shape[1].
        # Checked only for host-resident lengths (a device tensor would need a sync).
        if (src_tokens.dtype.is_floating_point or src_tokens.dtype == torch.int16) and src_lengths.device.type == "cpu" \
                and src_lengths.numel():
            if int(src_lengths.max()) != src_tokens.shape[1]:
                raise ValueError("src_tokens must be padded to max(src_lengths) exactly (got L=%d, max length %d)"
                                 % (src_tokens.shape[1], int(src_lengths.max())))

    @torch.no_grad()
    def _get_w2v_feature(self, src_tokens, src_lengths):
        self._check_inputs(src_tokens, src_lengths)
        B, L = src_tokens.shape
        plan = self._plan(B, L)
        plan.load_inputs(src_tokens if src_tokens.dtype == torch.int16 else src_tokens.float(), src_lengths)
        self.last_launches = plan.run(upto="w2v")
        return plan.view("w2v_out").clone(), plan.view("frame_mask"), plan.w2v_len64.clone()

    @torch.no_grad()
    def forward(self, src_tokens, src_lengths, **extra_args):      # extra: the collater's stray `mask=` kwarg
        self._check_inputs(src_tokens, src_lengths, allow_text=True)
        B, L = src_tokens.shape
        if not src_tokens.dtype.is_floating_point and src_tokens.dtype != torch.int16:     # text (MT) branch, interlingua:212-217
            if self.no_interlingua:
                raise NotImplementedError("no_interlingua with text input")
            plan = self._plan(B, L, text=True)
            plan.load_inputs(src_tokens.long(), src_lengths)
            self.last_launches = plan.run()
            out = plan.memories().to(self.encoder_out_dtype or torch.float32).clone(memory_format=torch.contiguous_format)
            return EncoderOut(out, torch.zeros(B, out.shape[0], dtype=torch.bool, device=out.device), None, None, None, None)
        plan = self._plan(B, L)
        plan.load_inputs(src_tokens if src_tokens.dtype == torch.int16 else src_tokens.float(), src_lengths)
        self.last_launches = plan.run()
        if self.base_encoder:                                      # w2v2_transformer.py:364-386
            out = plan.view("h_enc").transpose(0, 1)
            out = out.to(self.encoder_out_dtype or self._out_dtype(src_tokens)).clone(memory_format=torch.contiguous_format)
            T2 = out.shape[0]
            pad = torch.arange(T2, device=out.device)[None, :] >= plan.sub_valid[:B, None]
            return EncoderOut(out, pad if bool(pad.any()) else None, None, None, None, None)
        if self.no_interlingua:                                    # interlingua:260-262
            out = plan.view("h_enc").transpose(0, 1)
        else:
            out = plan.memories()
        # always a fresh tensor: for B == 1 `.contiguous()` would return a view of the plan's arena
        out = out.to(self.encoder_out_dtype or self._out_dtype(src_tokens)).clone(memory_format=torch.contiguous_format)
        pad = torch.zeros(B, out.shape[0], dtype=torch.bool, device=out.device)
        return EncoderOut(out, pad, None, None, None, None)


def build_encoder_from_state_dict(state_dict, interlingua_length=None, dtype=torch.float32, device="cuda",
                                  use_graph=True, conv_fp16=None):
    """Convenience: infer M from the checkpoint, load strictly, move to the device."""
    sd = state_dict
    if any(k.startswith("encoder.") for k in sd):
        sd = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    M = interlingua_length or sd["interlingua_embedding.weight"].shape[0]
    dead = "wav2vec_model.mask_emb" in sd
    vocab = sd["text_embed_tokens.weight"].shape[0] if "text_embed_tokens.weight" in sd else 0
    n_excl = 0
    while f"audio_exclusive_layers.{n_excl}.fc1.weight" in sd:
        n_excl += 1
    enc = B200InterlinguaEncoder(M, dtype=dtype, use_graph=use_graph, dead_heads=dead, text_vocab=vocab, conv_fp16=conv_fp16,
                                 modal_embedding="modal_embedding.weight" in sd, non_shared_encoder_layers=n_excl)
    enc.load_state_dict(sd, strict=True)
    return enc.to(device).eval()
