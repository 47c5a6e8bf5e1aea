#!/usr/bin/env python
"""Benchmark of the Chimera-ST speech-encoding hot path (waveform -> M shared semantic memories).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--dtype bf16|fp32]

Metric (BASELINE.json): encoded audio-seconds per second (valid audio only, padding not counted).
Workload `c3` (BASELINE.json configs[2]): Chimera-16, `--utts` utterances of U{2..30} s per GPU, sorted by
length, token-budget batches (max_tokens 2 000 000 samples, batch-size multiple 8: the reference's
batch_by_size rule), every rank works on its own utterance set (weak scaling, no data-path collective).
One "step" = one pass of the hot path over all of a rank's batches.

Printed (rank 0): ONE JSON line with value (inputs resident in HBM), e2e (pinned host buffers -> H2D ->
encoder.forward -> D2H of the memories, through the public module API), roofline of the dominant kernel
(live CUDA-event timing of every launch of it in one instrumented pass), cpu_baseline (the UNMODIFIED reference
encoder -- `oracle/_ref/src`, the bundle `oracle/make_ref_bundle.py` ships -- on this box's cores, one batch per length
decile of the workload), parity (the B200 memories of those same batches against the reference's, every row), clocks
sampled during the timed region.
Also in the line: `roofline.large_m` (the GEMM on the M >= 16 384 shapes), `roofline_hbm_kernel` (conv0 + GroupNorm + GELU), `e2e.int16_wire`
(the same end-to-end loop with 16-bit PCM on the wire) and, for the default workload, `secondary`: brief measurements of BASELINE configs[3]
(`c4_encode_decode`: Chimera-64, 64 x 20 s, encode + greedy decode, serial and pipelined over stream lanes), configs[4] (`c5_train`: the
CUDA-graphed training step with the bucketed gradient all-reduce, with and without the fused Adam update) and, for N > 1, a strong-scaling
probe (`c3_strong_scaling`).  `--workload c5` prints the training step as its own line (per-kernel split, ST + MT + contrastive form).
`--impl reference`: the reference arm = the reference's own modules (kind "reference"; only if the bundle is missing the
oracle port, kind "port", with a warning) on the host cores, same metric/config; step i times decile batch i mod 10.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import chimera_st_b200  # noqa: E402,F401
from chimera_st_b200 import synth, batching  # noqa: E402,F401
from chimera_st_b200 import distributed as D  # noqa: E402

SR = 16000


# --------------------------------------------------------------------------------------- workload
def make_workload(name, rank, world, utts, max_tokens=2000000, strong=False):
    """-> this rank's list of batches, each a list of utterance lengths (samples), longest first.
    c3: ONE global set of utts*world utterances, batched globally with the reference's rule and dealt
    round-robin to the ranks (generate.py:145-160 / ShardedIterator): per-GPU work stays ~constant as the
    number of GPUs grows (weak scaling) and the batch -> rank map is deterministic."""
    if name == "c3":
        rng = np.random.RandomState(2024)
        # weak scaling: utts per GPU (set grows with the ranks); strong scaling: a FIXED global set of `utts` utterances
        lens = rng.randint(32000, 480000 + 1, size=utts if strong else utts * world).astype(np.int64)
        mine, _ = D.shard_utterances(lens, world, rank, max_tokens, 8)
        return [[int(lens[i]) for i in b] for b in mine]
    if name == "c1":
        return [[80000] * 4]
    if name == "c2":
        return [[240000] * 32]
    if name == "c4":
        return [[320000] * 64]
    raise ValueError(name)


def host_batch(lens, seed):
    """Synthetic padded batch in pinned host memory (clamp(0.1*randn), zero tail)."""
    L = max(lens)
    g = torch.Generator().manual_seed(seed)
    x = torch.empty(len(lens), L, dtype=torch.float32)
    x.normal_(0.0, 0.1, generator=g).clamp_(-1.0, 1.0)
    for b, n in enumerate(lens):
        x[b, n:] = 0.0
    if torch.cuda.is_available():
        x = x.pin_memory()
    return x, torch.tensor(lens, dtype=torch.int64)


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------- kernel profiler
class LaunchProfiler:
    """CUDA-event bracket around every C-ABI call of one instrumented (eager, non-graph) pass: a proxy of the
    ctypes library records (kernel family, algorithmic flop, algorithmic bytes, start/end events)."""

    def __init__(self, lib):
        self.lib, self.rec = lib, []

    def __getattr__(self, name):
        fn = getattr(self.lib, name)
        if not name.startswith("cst_") or name.endswith("_ws_bytes") or name in ("cst_last_error", "cst_abi_version", "cst_device_info"):
            return fn

        def timed(*a):
            s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
            s.record()
            rc = fn(*a)
            e.record()
            self.rec.append(self._describe(name, a) + (s, e))
            return rc
        return timed

    @staticmethod
    def _describe(name, a):
        if name == "cst_gemm":
            p = a[0]._obj
            nz = p.nb_outer * p.nb_inner
            es = 2 if p.ab_dtype in (1, 2) else 4
            cs = 2 if p.c_dtype in (1, 2) else 4
            n_out = p.N // 2 if p.act == 3 else p.N
            by = (p.M * min(p.K, p.lda) + p.N * p.K) * es * nz + p.M * n_out * cs * nz + (p.M * n_out * 4 * nz if p.residual else 0)
            return ("gemm_tc_bf16" if p.ab_dtype in (1, 2) else "gemm_ffma_f32", 2.0 * p.M * p.N * p.K * nz, by,
                    "M%d N%d K%d z%d act%d" % (p.M, p.N, p.K, nz, p.act))
        if name == "cst_attention":
            dtype, B, H, n_q, n_kv = a[4], a[8], a[9], a[10], a[12]
            kind = "attention_tc_bf16" if (dtype == 1 and n_q > 64) else ("memory_attention" if (n_q <= 64 and n_kv <= 2048) else "attention_simt")
            return (kind, 4.0 * B * H * n_q * n_kv * 64, 0, "B%d H%d q%d kv%d" % (B, H, n_q, n_kv))
        if name == "cst_attention_segs":
            from chimera_st_b200 import plan as _plan_mod
            dtype, B, H, n_q, n_kv = a[4], a[8], a[9], a[11], a[12]
            kind = "attention_tc_bf16" if (dtype == 1 and n_q > 64) else ("memory_attention" if (n_q <= 64 and n_kv <= 2048) else "attention_simt")
            return (kind, 4.0 * H * _plan_mod.SEG_WORK.get(a[10], B * n_q * n_kv) * 64, 0, "B%d H%d q<=%d kv<=%d" % (B, H, n_q, n_kv))
        if name == "cst_layernorm":
            rows, C = a[8], a[9]
            by = rows * C * (4 + (4 if a[4] else 0) + ((2 if a[6] == 1 else 4) if a[5] else 0))
            return ("layernorm", 0.0, by, "rows%d C%d" % (rows, C))
        if name in ("cst_conv0_apply", "cst_conv0_apply_tc"):
            B, L, odt, rps = a[1], a[2], a[6], a[7]
            return ("conv0_gn_gelu", 2.0 * B * ((L - 10) // 5 + 1) * 512 * 10, B * rps * 512 * (2 if odt in (1, 2) else 4) + 4 * B * L,
                    "B%d L%d" % (B, L))
        if name == "cst_conv0_stats":
            return ("conv0_stats", 0.0, 4 * a[1] * a[2], "B%d L%d" % (a[1], a[2]))
        if name in ("cst_posconv", "cst_posconv_stacked"):
            B, n = a[5], a[6]
            return ("posconv_tc_bf16", 2.0 * B * n * 768 * 48 * 128, 0, "B%d T%d" % (B, n))
        return (name[4:], 0.0, 0, "")

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for kind, fl, by, _, s, e in self.rec:
            a = agg.setdefault(kind, [0, 0.0, 0.0, 0.0])
            a[0] += 1; a[1] += fl; a[2] += by; a[3] += s.elapsed_time(e) * 1e-3
        return {k: {"launches": v[0], "flops": v[1], "bytes": v[2], "seconds": v[3]} for k, v in agg.items()}


# --------------------------------------------------------------------------------------- arms
def reference_forward(M, device="cpu", dtype=torch.float32):
    """-> (fn(wave, lens) -> memories [M,B,512], kind).  kind "reference": the UNMODIFIED reference encoder
    (S2T_W2V2_TransformerInterlinguaEncoder, fairseq/models/chimera/w2v2_transformer_interlingua.py:155-312) imported
    from /root/reference or from the bundle oracle/_ref/src; kind "port": the oracle restatement (only when no
    reference tree is available -- says so on stderr)."""
    from oracle import make_overlay
    sd = synth.make_state_dict(seed=0, interlingua_length=M)
    if make_overlay.available():
        import warnings
        warnings.filterwarnings("ignore")
        from oracle.ref_model import build_reference_encoder
        enc, _ = build_reference_encoder(M)
        enc.load_state_dict(sd, strict=True)
        enc = enc.to(device=device, dtype=dtype).eval()

        def fn(wave, lens):
            return enc(wave.to(dtype), lens).encoder_out
        return fn, "reference"
    print("WARNING: no reference tree (/root/reference or oracle/_ref/src): timing the oracle PORT instead "
          "(run `python -m oracle.make_ref_bundle` in the dev container)", file=sys.stderr)
    from oracle import chimera_oracle as O
    sd = {k: v.to(device) for k, v in sd.items()}

    def fn(wave, lens):
        return O.encoder_forward(sd, wave, lens)[0]
    return fn, "port"


def decile_sample(batches, workload):
    """Bounded sample of the workload for the CPU arm: indices of one WHOLE bucketed batch per length decile (c3), or an
    8-utterance slice of the single fixed batch (c1/c2/c4).  -> ([(batch index, row count)], description)"""
    if workload != "c3" or len(batches) <= 10:
        rows = min(8, len(batches[0]))
        return [(0, rows)], "rows 0..%d of the %d x %d-sample batch" % (rows - 1, len(batches[0]), max(batches[0]))
    order = sorted(range(len(batches)), key=lambda i: max(batches[i]))
    pick = [order[min(len(order) - 1, (2 * d + 1) * len(order) // 20)] for d in range(10)]
    desc = "one whole bucketed batch per length decile: " + ", ".join(
        "%dx%d" % (len(batches[i]), max(batches[i])) for i in pick)
    return [(i, len(batches[i])) for i in pick], desc


def eager_gpu_rate(fn, wave, tl, steps, warmup, autocast=False):
    """BASELINE.md §5 comparison point: the reference's own modules as PyTorch-eager kernels (cuDNN conv1d, cuBLAS, SDPA,
    ATen norms / GELU) on the SAME GPU, one padded batch, CUDA events.  Not the reference arm the driver times (that is
    the CPU path); reported by `--impl reference --ref-device cuda`."""
    times = []
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        for i in range(warmup + steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(wave, tl)
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                times.append(e0.elapsed_time(e1) * 1e-3)
    return times



# --------------------------------------------------------------------------------------- secondary configurations in the default line
def c5_quick(rank, world, steps=5, dtype=torch.bfloat16, comm_dtype=torch.bfloat16, bucket_mb=32):
    """BASELINE configs[4] in brief (all ranks call it): the CUDA-graphed training step of `run_c5` (B = 8 x 150 000 samples per GPU,
    contrastive head, bucketed gradient all-reduce overlapped with the backward segments), timed with CUDA events, max over ranks."""
    from chimera_st_b200.train import EncoderTrainStep, GraphedTrainStep
    from chimera_st_b200 import ddp, losses
    B, Lw, M = C5_B, C5_L, C5_M
    sd = synth.make_state_dict(seed=0, interlingua_length=M, dead_heads=False)
    step = EncoderTrainStep(sd, B, Lw, device="cuda", feature_grad_mult=0.1, dtype=dtype)
    host_w, host_l = host_batch([Lw] * B, seed=1000 * rank)
    wave, lens = host_w.cuda(), host_l.cuda()
    text_mem = torch.randn(M, B, 512, generator=torch.Generator().manual_seed(7 + rank)).cuda()

    def loss_fn(mem):
        _, loss, da, _ = losses.contrastive_loss(mem.contiguous(), text_mem, temp=0.1, grad_scale=1.0)
        return loss, da
    gs = GraphedTrainStep(step, wave, lens, loss_fn)
    red = ddp.GradAllReducer(gs.names, world_size=world, bucket_bytes=bucket_mb << 20, comm_dtype=comm_dtype)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}
    for key, reducer in (("ms_per_step", red), ("ms_per_step_without_allreduce", None)):
        gs.reducer = reducer
        for _ in range(3):
            gs.run()
        torch.cuda.synchronize(); D.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            gs.run()
        e1.record()
        torch.cuda.synchronize(); D.barrier()
        out[key] = round(D.reduce_max(e0.elapsed_time(e1), "cuda") / steps, 3)
    total_audio = D.reduce_sum(B * Lw / SR, "cuda")
    # ... and with the parameter update (fused Adam on the reduced gradients + operand refresh, one more CUDA graph)
    from chimera_st_b200.train import FusedAdam
    gs.reducer = ddp.GradAllReducer(gs.names, world_size=world, bucket_bytes=bucket_mb << 20, comm_dtype=comm_dtype, persistent=True)
    gs.optimizer = FusedAdam({n: step.sd[n] for n, _ in gs.names}, lr=1e-5, betas=(0.9, 0.98), eps=1e-8)
    loss0 = float(gs.loss)
    for _ in range(3):
        gs.run()
    torch.cuda.synchronize(); D.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        gs.run()
    e1.record()
    torch.cuda.synchronize(); D.barrier()
    t_opt = D.reduce_max(e0.elapsed_time(e1), "cuda") / steps
    out["with_adam_update"] = {"ms_per_step": round(t_opt, 3), "audio_s_per_s": round(total_audio / (t_opt * 1e-3), 1),
                               "loss_before": loss0, "loss_after_%d_updates" % (3 + steps): float(gs.loss)}
    # ... and the same step with the recipe's dropout (p = 0.1 at every site incl. the attention probabilities; masks regenerated, not stored)
    final_loss, n_grad = float(gs.loss), sum(n for _, n in gs.names)
    try:
        gs = None
        step_d = EncoderTrainStep(sd, B, Lw, device="cuda", feature_grad_mult=0.1, dtype=dtype, dropout=0.1, w2v_dropout=0.1,
                                  w2v_dropout_input=0.1, seed=1 + rank)
        gs = GraphedTrainStep(step_d, wave, lens, loss_fn)
        gs.reducer = ddp.GradAllReducer(gs.names, world_size=world, bucket_bytes=bucket_mb << 20, comm_dtype=comm_dtype)
        for _ in range(3):
            gs.run()
        torch.cuda.synchronize(); D.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            gs.run()
        e1.record()
        torch.cuda.synchronize(); D.barrier()
        t_d = D.reduce_max(e0.elapsed_time(e1), "cuda") / steps
        out["with_dropout_0.1"] = {"ms_per_step": round(t_d, 3), "audio_s_per_s": round(total_audio / (t_d * 1e-3), 1),
                                   "sites": len(step_d._sites), "note": "elementwise sites + attention probabilities (materialised attention forward)"}
        del step_d
    except Exception as e:                                         # a secondary figure must not take the headline line down
        out["with_dropout_0.1"] = {"error": repr(e)[:200]}
    out.update({"workload": "c5: training step (forward + contrastive head + backward + gradient all-reduce), B=%d x %d samples per GPU, bf16" % (B, Lw),
                "audio_s_per_s": round(total_audio / (out["ms_per_step"] * 1e-3), 1), "n_gpus": world,
                "allreduce_bytes_per_step": n_grad * (2 if comm_dtype is not None else 4),
                "allreduce_exposed_ms": round(out["ms_per_step"] - out["ms_per_step_without_allreduce"], 3), "loss": final_loss})
    del gs, step
    torch.cuda.empty_cache()
    return out


def c4_decode_quick(rank, world, dtype=torch.bfloat16, beam=1):
    """BASELINE configs[3] in brief (all ranks, one batch each): Chimera-64, 64 utterances x 20 s, encode + greedy decode
    (SequenceGenerator beam 1, max_len_b 200), CUDA events, max over ranks."""
    from chimera_st_b200.encoder import build_encoder_from_state_dict
    from chimera_st_b200.decoder import B200GreedyDecoder, B200BeamDecoder
    M, lens_ = 64, [320000] * 64
    enc = build_encoder_from_state_dict(synth.make_state_dict(seed=0, interlingua_length=M), dtype=dtype, device="cuda", use_graph=True)
    dsd = synth.make_decoder_state_dict(seed=1)
    dec = B200GreedyDecoder(dsd, dtype=dtype, device="cuda") if beam == 1 else B200BeamDecoder(dsd, beam=beam, dtype=dtype, device="cuda")
    w, l = host_batch(lens_, seed=77 + rank)
    w, l = w.cuda(), l.cuda()
    for _ in range(2):
        dec.generate(enc(w, l).encoder_out, max_len=200)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize(); D.barrier(); torch.cuda.synchronize()
    ev[0].record()
    mem0 = enc(w, l).encoder_out
    ev[1].record()
    dec.generate(mem0, max_len=200)
    ev[2].record()
    torch.cuda.synchronize()
    t_enc, t_dec = D.reduce_max(ev[0].elapsed_time(ev[1]), "cuda"), D.reduce_max(ev[1].elapsed_time(ev[2]), "cuda")
    audio = D.reduce_sum(sum(lens_) / SR, "cuda")
    pipe = None
    if beam == 1:
        # throughput form over a stream of batches: the latency-bound decode of batch i runs on its own stream underneath the
        # encoder pass of batch i+1 and the decodes of batches i-1, i-2 (decoder.generate_async); same kernels, same hypotheses
        n_lanes = int(os.environ.get("CST_C4_LANES", "6"))
        n_batches = 3 * n_lanes
        pending = [None] * n_lanes

        def run_pipeline(k):
            for i in range(k):
                ln = i % n_lanes
                if pending[ln] is not None:
                    dec.collect(pending[ln])
                pending[ln] = dec.generate_async(enc(w, l).encoder_out, max_len=200, lane=ln)
            for ln in range(n_lanes):
                if pending[ln] is not None:
                    dec.collect(pending[ln])
                    pending[ln] = None
        run_pipeline(n_lanes)                                    # graph capture / warm-up of every lane
        torch.cuda.synchronize(); D.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev[0].record()
        run_pipeline(n_batches)
        ev[1].record()
        torch.cuda.synchronize()
        t_pipe = D.reduce_max(ev[0].elapsed_time(ev[1]), "cuda")
        pipe = {"lanes": n_lanes, "batches": n_batches, "ms_per_batch": round(t_pipe / n_batches, 3),
                "encode_decode_audio_s_per_s": round(audio * n_batches / (t_pipe * 1e-3), 1), "wall_s": round(time.perf_counter() - t0, 3)}
    out = {"workload": "c4: Chimera-64 encode + %s decode, 64 utterances x 20 s per GPU, bf16" % ("greedy" if beam == 1 else "beam-%d" % beam),
           "pipelined": pipe,
           "n_gpus": world, "encode_ms": round(t_enc, 3), "decode_ms": round(t_dec, 3), "decode_steps": dec.last_steps,
           "us_per_decode_step": round(1e3 * t_dec / max(1, dec.last_steps), 1),
           "encode_decode_audio_s_per_s": round(audio / ((t_enc + t_dec) * 1e-3), 1),
           "encode_only_audio_s_per_s": round(audio / (t_enc * 1e-3), 1)}
    del enc, dec
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------- c5: training step
C5_B, C5_L, C5_M = 8, 150000, 16            # ~1.2 M samples per GPU, T' = 468 frames per utterance


def run_c5(args, rank, world, local_rank, cores):
    """BASELINE configs[4]: Chimera-16 ST training forward + backward of the encoder / memory path, ~1.2 M audio samples per GPU
    (B = 8 x 150 000), data-parallel over the ranks with a bucketed NCCL all-reduce of the gradients overlapped with the backward
    segments (`chimera_st_b200.ddp`, replacing LegacyDistributedDataParallel).  The loss is the contrastive (InfoNCE) head over the
    memories against fixed synthetic text-pass memories (triplet_st_mt_contrastive.py:154-169); dropout 0.1 at every site of the recipe
    (--dropout 0 = the parity configuration), LayerDrop 0 (it needs one CUDA graph per pattern: eager path only).
    step = forward + loss + backward + gradient all-reduce;  value: waveforms resident in HBM;  e2e: pinned host waveforms -> H2D ->
    step -> D2H of the loss."""
    B, Lw, M = C5_B, C5_L, C5_M
    audio_per_step = B * Lw / SR
    dt_name = args.dtype
    cfg = {"workload": "c5: Chimera-16 encoder+memory TRAINING step (forward + contrastive head + backward + gradient all-reduce), "
                       "B=%d x %d samples (%.1f M samples, %.1f audio-s) per GPU" % (B, Lw, B * Lw / 1e6, audio_per_step),
           "interlingua_length": M, "parallelism": "data-parallel dp%d, bucketed gradient all-reduce (%s buckets of %d MB) overlapped with "
                                                   "the backward segments" % (world, args.comm_dtype, args.bucket_mb),
           "dropout": (0.0 if not getattr(args, "dropout", 0.0) else
                       "%.2f at every site of the recipe: elementwise dropouts, dropout_input and the attention probabilities (Philox masks "
                       "regenerated in the backward pass, never stored; layers with attention dropout run the materialised attention on "
                       "batched tcgen05 GEMMs instead of the flash kernel)" % args.dropout),
           "layerdrop": 0.0, "feature_grad_mult": 0.1,
           "l2": "tape (> 4 GB of activations per step) exceeds the 126 MB L2; no explicit flush",
           "precision": ("bf16 GEMM operands (fp16 in the conv-stack forward), fp32 accumulation / gradients / norms / softmax"
                         if dt_name == "bf16" else "fp32 FFMA")}
    if args.impl == "reference":
        if rank != 0:
            return
        # the reference differentiates its modules with autograd; its own forward cannot be differentiated under torch 2.11 (in-place
        # masking on a view), so the CPU arm is autograd through the oracle restatement of the same modules
        from oracle import chimera_oracle as O
        torch.set_num_threads(cores)
        sd = synth.make_state_dict(seed=0, interlingua_length=M, dead_heads=False)
        rows = 2                                               # bounded sample: 2 of the 8 utterances per step
        wave, tl = host_batch([Lw] * rows, seed=0)
        R = torch.randn(M, rows, 512, generator=torch.Generator().manual_seed(1))
        times = []
        for it in range(args.warmup + args.steps):
            sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() else v) for k, v in sd.items()}
            t = time.perf_counter()
            mem, _ = O.encoder_forward(sdg, wave, tl)
            (mem * R).sum().backward()
            dtm = time.perf_counter() - t
            if it >= args.warmup:
                times.append(dtm)
        rate = rows * Lw / SR * len(times) / sum(times)
        sample = "%d of the %d utterances of the step (L=%d), forward + backward by autograd, fp32, torch CPU %d threads" % (rows, B, Lw, cores)
        print(json.dumps({"impl": "reference", "metric": "encoded audio-sec/sec", "value": round(rate, 3), "unit": "audio-s/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * sum(times) / len(times), 2),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": round(rate, 3), "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": round(rate, 3), "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    D.init("nccl")
    from chimera_st_b200.train import EncoderTrainStep, GraphedTrainStep
    from chimera_st_b200 import ddp, losses
    dtype = torch.bfloat16 if dt_name == "bf16" else torch.float32
    sd = synth.make_state_dict(seed=0, interlingua_length=M, dead_heads=False)
    pdrop = float(getattr(args, "dropout", 0.0))
    step = EncoderTrainStep(sd, B, Lw, device="cuda", feature_grad_mult=0.1, dtype=dtype, dropout=pdrop, w2v_dropout=pdrop,
                            w2v_dropout_input=pdrop, seed=1 + rank)          # every rank draws its own masks
    host_w, host_l = host_batch([Lw] * B, seed=1000 * rank)
    wave, lens = host_w.cuda(), host_l.cuda()
    text_mem = torch.randn(M, B, 512, generator=torch.Generator().manual_seed(7 + rank)).cuda()

    def loss_fn(mem):
        _, loss, da, _ = losses.contrastive_loss(mem.contiguous(), text_mem, temp=0.1, grad_scale=1.0)
        return loss, da

    gs = GraphedTrainStep(step, wave, lens, loss_fn)
    comm = torch.bfloat16 if args.comm_dtype == "bf16" else None
    gs.reducer = ddp.GradAllReducer(gs.names, world_size=world, bucket_bytes=args.bucket_mb << 20, comm_dtype=comm)
    n_params = sum(n for _, n in gs.names)
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        D.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return gs.run()

    def step_e2e():
        wave.copy_(host_w, non_blocking=True)
        lens.copy_(host_l, non_blocking=True)
        loss = gs.run()
        loss_host.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    t_res = D.reduce_max(e0.elapsed_time(e1) * 1e-3, "cuda")
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    t_e2e = D.reduce_max(e0.elapsed_time(e1) * 1e-3, "cuda")
    # local compute only (no all-reduce): what the collective costs on top
    red, gs.reducer = gs.reducer, None
    barrier()
    e0.record()
    for _ in range(args.steps):
        gs.run()
    e1.record()
    barrier()
    t_local = D.reduce_max(e0.elapsed_time(e1) * 1e-3, "cuda")
    gs.reducer = red
    # replicas agree after the all-reduce (every rank holds the same averaged gradient)
    gs.run()
    torch.cuda.synchronize()
    probe = gs.reduced["wav2vec_model.encoder.layers.0.fc1.weight"].float()
    csum = float(probe.double().sum())
    agree = abs(D.reduce_max(csum, "cuda") + D.reduce_max(-csum, "cuda")) <= 1e-6 * max(1.0, abs(csum))
    total_audio = D.reduce_sum(audio_per_step, "cuda")
    # the same step with the parameter update: fused Adam on the reduced gradients + refresh of the kernel-layout operands, one more
    # CUDA graph (the reference: FP16Optimizer + Adam, fairseq/optim/fp16_optimizer.py, adam.py)
    from chimera_st_b200.train import FusedAdam
    gs.reducer = ddp.GradAllReducer(gs.names, world_size=world, bucket_bytes=args.bucket_mb << 20, comm_dtype=comm, persistent=True)
    gs.optimizer = FusedAdam({n: step.sd[n] for n, _ in gs.names}, lr=1e-5, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.0)
    loss0 = float(gs.loss)
    for _ in range(3):
        gs.run()
    barrier()
    e0.record()
    for _ in range(args.steps):
        gs.run()
    e1.record()
    barrier()
    t_opt = D.reduce_max(e0.elapsed_time(e1) * 1e-3, "cuda")
    loss1 = float(gs.loss)
    psum = float(step.sd["wav2vec_model.encoder.layers.0.fc1.weight"].double().sum())
    params_agree = abs(D.reduce_max(psum, "cuda") + D.reduce_max(-psum, "cuda")) <= 1e-9 * max(1.0, abs(psum))
    gs.optimizer = None
    # the ST + MT + contrastive form of the step (criterions/triplet_st_mt_contrastive.py): a text pass of the SAME encoder (8 x 40
    # tokens) inside the forward graph, InfoNCE between the audio and the text memories, both backward passes into one gradient set
    joint = None
    try:
        from chimera_st_b200.train import TextTrainPass
        sd_t = synth.make_state_dict(seed=0, interlingua_length=M, dead_heads=False, text_vocab=10000)
        step_t = EncoderTrainStep(sd_t, B, Lw, device="cuda", feature_grad_mult=0.1, dtype=dtype)
        tp = TextTrainPass(step_t, B, 40)
        tok = torch.randint(4, 10000, (B, 40), generator=torch.Generator().manual_seed(11 + rank)).cuda()
        tlen = torch.full((B,), 40, dtype=torch.int64).cuda()

        def joint_loss(mem):
            tmem = tp.forward(tok, tlen)
            _, loss, da, dt_ = losses.contrastive_loss(mem.contiguous(), tmem.contiguous(), temp=0.1, grad_scale=1.0)
            return loss, da, tp.backward(dt_)
        gj = GraphedTrainStep(step_t, wave, lens, joint_loss)
        gj.reducer = ddp.GradAllReducer(gj.names, world_size=world, bucket_bytes=args.bucket_mb << 20, comm_dtype=comm)
        for _ in range(3):
            gj.run()
        barrier()
        e0.record()
        for _ in range(args.steps):
            gj.run()
        e1.record()
        barrier()
        t_j = D.reduce_max(e0.elapsed_time(e1) * 1e-3, "cuda")
        joint = {"ms_per_step": round(1e3 * t_j / args.steps, 3), "text_tokens_per_gpu": B * 40, "loss": float(gj.loss),
                 "gradient_tensors": len(gj.names), "parameters": sum(n for _, n in gj.names)}
        del gj, tp, step_t
    except Exception as e:                                            # noqa: BLE001
        joint = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}

    # ---- instrumented eager pass: per-kernel CUDA-event timing for the roofline object
    prof = LaunchProfiler(step.o.lib)
    step.o.lib = prof
    torch.cuda._sleep(int(2e8))
    loss, dmem = loss_fn(step.forward(wave, lens))
    step.backward(dmem)
    step.o.lib = prof.lib
    ksum = prof.summary()
    launches_per_step = sum(v["launches"] for v in ksum.values())
    if rank != 0:
        D.finalize()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    dom = "gemm_tc_bf16" if "gemm_tc_bf16" in ksum else "gemm_ffma_f32"
    kd = ksum[dom]
    ach = kd["flops"] / kd["seconds"] / 1e12
    peak = peaks.get("bf16_tflops_sustained", 1400.0) if dom == "gemm_tc_bf16" else 72.0
    step_kernel_s = sum(v["seconds"] for v in ksum.values())
    roofline = {"bound": "tensor", "kernel": dom, "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4),
                "traffic": None, "peak_source": "measured (sustained)" if peaks and dom == "gemm_tc_bf16" else "nominal",
                "launches_per_step": kd["launches"], "share_of_kernel_time": round(kd["seconds"] / step_kernel_s, 3),
                "algorithmic_flop_per_step": kd["flops"],
                "by_kernel": {k: {"launches": v["launches"], "ms": round(v["seconds"] * 1e3, 3),
                                  "tflops": round(v["flops"] / v["seconds"] / 1e12, 2) if v["flops"] > 0 else None}
                              for k, v in sorted(ksum.items(), key=lambda kv: -kv[1]["seconds"])}}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import chimera_oracle as O
        torch.set_num_threads(cores)
        rows = 2
        w2, l2 = host_w[:rows].clone(), host_l[:rows].clone()
        R = torch.randn(M, rows, 512, generator=torch.Generator().manual_seed(1))
        sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() else v) for k, v in sd.items()}
        t = time.perf_counter()
        mem, _ = O.encoder_forward(sdg, w2, l2)
        (mem * R).sum().backward()
        t_cpu = time.perf_counter() - t
        cpu = {"value": round(rows * Lw / SR / t_cpu, 3), "unit": "audio-s/s", "cores": cores, "kind": "port",
               "sample": "%d of the %d utterances (L=%d), forward + backward by autograd through the oracle, fp32, torch CPU %d threads, "
                         "one pass" % (rows, B, Lw, cores)}
    grad_bytes = n_params * (2 if comm is not None else 4)
    line = {"metric": "encoded audio-sec/sec", "value": round(total_audio * args.steps / t_res, 1), "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(1e3 * t_res / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if dtype == torch.bfloat16 else "f32",
            "data": "synthetic", "config": cfg,
            "e2e": {"value": round(total_audio * args.steps / t_e2e, 1), "unit": "audio-s/s",
                    "h2d_bytes_per_step": host_w.numel() * 4 + host_l.numel() * 8, "d2h_bytes_per_step": 4},
            "gpu_launches": launches_per_step * args.steps, "cuda_graph": True,
            "train": {"loss": float(gs.loss), "parameters": n_params, "gradient_tensors": len(gs.names),
                      "backward_segments": len(gs.seg_grads), "buckets": len(red.buckets),
                      "allreduce_bytes_per_step": grad_bytes, "ms_per_step_without_allreduce": round(1e3 * t_local / args.steps, 3),
                      "allreduce_exposed_ms": round(1e3 * (t_res - t_local) / args.steps, 3), "replicas_agree": bool(agree),
                      "with_adam_update": {"ms_per_step": round(1e3 * t_opt / args.steps, 3),
                                           "audio_s_per_s": round(total_audio * args.steps / t_opt, 1),
                                           "loss_before": loss0, "loss_after_%d_updates" % (3 + args.steps): loss1,
                                           "parameters_agree_across_ranks": bool(params_agree)},
                      "st_mt_contrastive_step": joint},
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line))
    D.finalize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c1", "c2", "c4", "c5"],
                    help="c5 = BASELINE configs[4]: training forward + backward of the path with the DDP gradient all-reduce")
    ap.add_argument("--comm-dtype", default="bf16", choices=["bf16", "fp32"], help="c5: dtype of the gradient buckets on the wire")
    ap.add_argument("--bucket-mb", type=int, default=32, help="c5: gradient bucket size")
    ap.add_argument("--dropout", type=float, default=0.1,
                    help="c5: probability of the recipe's dropouts (train-en2any-ST.sh:45 uses 0.1 = the timing configuration, SURVEY.md §7); "
                         "0 = the parity configuration")
    ap.add_argument("--utts", type=int, default=512)
    ap.add_argument("--max-tokens", type=int, default=2000000,
                    help="c3 token budget per batch in samples (default: the reference's --max-tokens 2000000, "
                         "chimera/scripts/interactive-en2any-ST.sh:21)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="c3 only: weak = --utts utterances per GPU (default, the driver's 1->8 run); strong = a fixed global set "
                         "of --utts utterances dealt over the ranks (exposes graph warm-up and tail imbalance)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--lanes", type=int, default=2, help="concurrent CUDA-stream lanes for independent super-batches")
    ap.add_argument("--super-rows", type=int, default=-1,
                    help="wav2vec2 frame rows per super-batch (several reference batches, each with its own padded width, in one "
                         "row space / one CUDA graph); -1 = the encoder's default (49152), 0 = one plan per reference batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="c3 only: skip the brief secondary measurements (c4 encode + greedy decode, c5 training step) appended to the line")
    ap.add_argument("--profile-json", default="")
    ap.add_argument("--beam", type=int, default=1, help="--decode only: beam width (1 = greedy, 2..8 = B200BeamDecoder)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: cpu = the reference arm (host cores); cuda = the same oracle port as "
                         "PyTorch-eager kernels on the GPU (BASELINE.md §5 comparison point), first batch of the workload")
    ap.add_argument("--decode", action="store_true",
                    help="also time greedy decoding (max_len_b 200) of the first batch's memories on the GPU decoder and "
                         "report it in a `decode` object (BASELINE configs[3]: encode + greedy decode); the headline "
                         "value / e2e stay the encoder path")
    args = ap.parse_args()

    rank, world, local_rank = D.env_rank_world()
    cores = os.cpu_count() or 1
    if args.workload == "c5":
        return run_c5(args, rank, world, local_rank, cores)
    M = 64 if args.workload == "c4" else 16          # configs[3] is Chimera-64
    scaling = args.scaling if args.workload == "c3" else "weak"
    batches = make_workload(args.workload, rank, world, args.utts, args.max_tokens, strong=scaling == "strong")
    audio_per_step = sum(sum(b) for b in batches) / SR
    cfg = {"workload": "%s: Chimera-%d encoder+memory, %s" % (args.workload, M, {
        "c3": ("%d utts/GPU U{2..30}s (global set x%d ranks, round-robin sharded)" % (args.utts, world) if scaling == "weak" else
               "FIXED global set of %d utts U{2..30}s dealt round-robin over %d ranks" % (args.utts, world)) +
              ", length-bucketed max_tokens=%.0e bsz%%8 (%d batches on rank 0)" % (args.max_tokens, len(batches)),
        "c1": "B=4 x 5 s", "c2": "B=32 x 15 s", "c4": "B=64 x 20 s"}[args.workload]),
        "interlingua_length": M, "batches_per_step": len(batches), "audio_sec_per_step_per_gpu": round(audio_per_step, 2),
        "parallelism": "utterance-sharded dp%d, no data-path collective" % world,
        "l2": "activations per batch (>=0.4 GB) exceed the 126 MB L2; no explicit flush",
        "precision": "bf16 operands / fp32 accumulate, residual stream, norms and softmax; conv feature extractor operands fp16 (same bytes)"
                     if args.dtype == "bf16" else "fp32 FFMA"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        if args.ref_device == "cuda":
            lens0 = batches[0]
            wave, tl = host_batch(lens0, seed=0)
            wave, tl = wave.cuda(), tl.cuda()
            res, kind = {}, None
            for name, dt, ac in (("fp32", torch.float32, False), ("fp16", torch.float16, False), ("bf16_autocast", torch.float32, True)):
                fn, kind = reference_forward(M, "cuda", dt)
                times = eager_gpu_rate(fn, wave, tl, max(1, args.steps), 2, ac)
                res[name] = {"audio_s_per_s": round(sum(lens0) / SR / statistics.median(times), 1),
                             "ms_per_batch": round(1e3 * statistics.median(times), 2)}
                del fn
                torch.cuda.empty_cache()
            print(json.dumps({"impl": "reference", "variant": "the reference's own modules as PyTorch-eager kernels on cuda "
                              "(kind %s; not the driver's reference arm)" % kind,
                              "metric": "encoded audio-sec/sec", "unit": "audio-s/s", "value": res["fp16"]["audio_s_per_s"],
                              "config": cfg, "batch": "first batch of the workload: %d utterances, L=%d" % (len(lens0), max(lens0)),
                              "eager_gpu": res, "device": torch.cuda.get_device_name(0)}))
            return
        torch.set_num_threads(cores)
        fn, kind = reference_forward(M)
        picks, desc = decile_sample(batches, args.workload)
        host = {i: host_batch(batches[i][:rows], seed=1000 * rank + i) for i, rows in picks}
        times, audio = [], []
        with torch.no_grad():
            for it in range(args.warmup + args.steps):
                i, rows = picks[it % len(picks)]
                w, l = host[i]
                t = time.perf_counter()
                fn(w, l)
                dt = time.perf_counter() - t
                if it >= args.warmup:
                    times.append(dt)
                    audio.append(sum(batches[i][:rows]) / SR)
        rate = sum(audio) / sum(times)
        sample = "%s; step i runs batch i mod %d (%d timed steps = %.0f audio-s), fp32, torch CPU %d threads" % (
            desc, len(picks), len(times), sum(audio), cores)
        line = {"impl": "reference", "metric": "encoded audio-sec/sec", "value": round(rate, 3), "unit": "audio-s/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(1e3 * sum(times) / len(times), 2), "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": round(rate, 3), "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": round(rate, 3), "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    D.init("nccl")
    from chimera_st_b200.encoder import build_encoder_from_state_dict
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    sd = synth.make_state_dict(seed=0, interlingua_length=M)
    enc = build_encoder_from_state_dict(sd, dtype=dtype, device="cuda", use_graph=not args.no_graph)
    enc.MAX_PLANS = 64          # bounded LRU of cached plans / CUDA graphs (c3: 19 super-batch compositions per lane set)

    host = [host_batch(b, seed=1000 * rank + i) for i, b in enumerate(batches)]
    dev = [(w.cuda(non_blocking=True), l.cuda(non_blocking=True)) for w, l in host]
    out_host = [torch.empty(M, len(b), 512, dtype=torch.float32).pin_memory() for b in batches]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        D.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return D.reduce_max(x, "cuda")

    lanes = max(1, args.lanes)
    super_rows = None if args.super_rows < 0 else args.super_rows
    supers = enc.plan_super_batches([tuple(w.shape) for w, _ in dev], super_rows)

    def step_resident():
        enc.forward_many(dev, n_lanes=lanes, super_rows=super_rows)
        return enc.last_launches

    def step_e2e():
        # public API with HOST buffers: pinned waveforms -> H2D -> encoder -> D2H of the memories
        # (the encoder copies host tensors straight into its input buffers on the stream that runs the batch, so
        # with stream lanes one batch's H2D copy overlaps another batch's kernels)
        enc.forward_many(host, n_lanes=lanes, out=out_host, super_rows=super_rows)
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    barrier()
    e0.record()
    for _ in range(args.steps):
        launches += step_resident()
    e1.record()
    barrier()
    t_res = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    t_e2e = max_over_ranks(max(e0.elapsed_time(e1) * 1e-3, 0.0))
    t_e2e_wall = time.perf_counter() - t0
    # the same end-to-end loop with 16-bit PCM on the wire (what a .wav holds; de-quantised on the GPU): half the H2D bytes
    host16 = [((w * 32768.0).round().clamp_(-32768, 32767).to(torch.int16).pin_memory(), l) for w, l in host]

    def step_e2e16():
        enc.forward_many(host16, n_lanes=lanes, out=out_host, super_rows=super_rows)
        torch.cuda.synchronize()
    for _ in range(2):
        step_e2e16()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_e2e16()
    e1.record()
    barrier()
    t_e2e16 = max_over_ranks(max(e0.elapsed_time(e1) * 1e-3, 0.0))
    del host16

    total_audio = D.reduce_sum(audio_per_step, "cuda")

    # ---- instrumented pass: per-kernel CUDA-event timing (non-graph) for the roofline object, same super-batches
    prof = None
    for idx in supers:
        groups = [tuple(dev[i][0].shape) for i in idx]
        p = enc._plan(groups[0][0], groups[0][1]) if len(groups) == 1 else enc._plan(None, None, groups=groups)
        if prof is None:
            prof = LaunchProfiler(p.lib)
        real = p.lib
        p.lib = prof
        for k, i in enumerate(idx):
            p.load_inputs(dev[i][0], dev[i][1], group=k)
        # keep the GPU busy while the host enqueues this super-batch's launches, so that the CUDA-event brackets
        # measure kernel execution and not host launch gaps (a graph replay has none either)
        torch.cuda._sleep(int(3.5e7))
        p.run(eager=True)
        p.lib = real
    ksum = prof.summary()

    # ---- secondary configurations (all ranks take part: c5 has the gradient all-reduce); never allowed to break the headline
    extras = None
    if args.workload == "c3" and not args.no_extras and dtype == torch.bfloat16:
        extras = {}
        if world > 1 and scaling == "weak":
            # strong-scaling probe next to the weak-scaling headline: ONE fixed global set of `--utts` utterances dealt over the
            # ranks (shows tail imbalance and per-rank launch floors the weak form hides); resident inputs, max over ranks
            try:
                sb = make_workload("c3", rank, world, args.utts, args.max_tokens, strong=True)
                sdev = [tuple(t.cuda() for t in host_batch(b, seed=5000 + 1000 * rank + i)) for i, b in enumerate(sb)]
                for _ in range(2):
                    enc.forward_many(sdev, n_lanes=lanes, super_rows=super_rows)
                barrier()
                e0.record()
                for _ in range(args.steps):
                    enc.forward_many(sdev, n_lanes=lanes, super_rows=super_rows)
                e1.record()
                barrier()
                t_s = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
                a_s = D.reduce_sum(sum(sum(b) for b in sb) / SR, "cuda")
                extras["c3_strong_scaling"] = {"workload": "c3, FIXED global set of %d utterances dealt round-robin over %d ranks" % (args.utts, world),
                                               "scaling": "strong", "n_gpus": world, "batches_on_rank0": len(sb),
                                               "ms_per_step": round(1e3 * t_s / args.steps, 3),
                                               "audio_s_per_s": round(a_s * args.steps / t_s, 1)}
                del sdev
            except Exception as e:                                    # noqa: BLE001
                extras["c3_strong_scaling"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
        enc.invalidate()
        torch.cuda.empty_cache()
        for name, fn in (("c4_encode_decode", lambda: c4_decode_quick(rank, world)), ("c5_train", lambda: c5_quick(rank, world))):
            try:
                extras[name] = fn()
            except Exception as e:                                    # noqa: BLE001
                extras[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}   # same code on every rank: ranks stay in step

    if rank != 0:
        D.finalize()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    dom = "gemm_tc_bf16" if "gemm_tc_bf16" in ksum else "gemm_ffma_f32"
    kd = ksum[dom]
    ach = kd["flops"] / kd["seconds"] / 1e12
    if dom == "gemm_tc_bf16":
        peak, peak_src = peaks.get("bf16_tflops_sustained", 1400.0), ("measured (sustained)" if peaks else "fallback")
    else:
        peak, peak_src = 72.0, "nominal fp32 FFMA 148 SM x 128 lanes x 2 x 1.9 GHz (no measured fp32 peak)"
    step_kernel_s = sum(v["seconds"] for v in ksum.values())
    # DRAM traffic per launch of the dominant kernel from the committed ncu capture (profiles/traffic_r02.json,
    # same workload family: c3 batches); null for other workloads
    traffic, traffic_src = None, None
    try:
        if args.workload == "c3":
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json")))
            traffic = round(tj["kernels"][dom]["traffic_bytes_per_launch"])
            traffic_src = "profiles/traffic_r02.json: " + tj["workload"]
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": dom, "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s",
                "frac": round(ach / peak, 4), "traffic": traffic, "traffic_unit": "bytes per launch (dram read+write, ncu)",
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": round(kd["bytes"] / max(1, kd["launches"])),
                "peak_source": peak_src,
                "launches_per_step": kd["launches"], "share_of_kernel_time": round(kd["seconds"] / step_kernel_s, 3),
                "by_kernel": {k: {"launches": v["launches"], "ms": round(v["seconds"] * 1e3, 3),
                                  "tflops": round(v["flops"] / v["seconds"] / 1e12, 2) if v["flops"] > 0 else None,
                                  "gbs": round(v["bytes"] / v["seconds"] / 1e9, 1) if v["bytes"] > 0 else None}
                              for k, v in ksum.items()}}
    # the same kernel on the wav2vec2-stage shapes only (M >= 16 384 rows): the small shared / memory-stage launches (M <= 7k rows,
    # a few CTAs wide) are 15 % of the GEMM time at a fraction of the rate and pull the family average down
    try:
        import re as _re
        big_fl = big_s = 0.0
        for kind, fl, by, desc, s_ev, e_ev in prof.rec:
            m_ = _re.match(r"M(\d+) ", desc) if kind == dom else None
            if m_ and int(m_.group(1)) >= 16384:
                big_fl += fl
                big_s += s_ev.elapsed_time(e_ev) * 1e-3
        if big_s > 0:
            roofline["large_m"] = {"min_rows": 16384, "achieved": round(big_fl / big_s / 1e12, 2), "frac": round(big_fl / big_s / 1e12 / peak, 4),
                                   "share_of_gemm_time": round(big_s / kd["seconds"], 3)}
    except Exception:
        pass
    hbm_roof = None
    if "conv0_gn_gelu" in ksum and ksum["conv0_gn_gelu"]["seconds"] > 0:
        k1 = ksum["conv0_gn_gelu"]
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        gbs = k1["bytes"] / k1["seconds"] / 1e9
        hbm_roof = {"bound": "hbm", "kernel": "conv0_gn_gelu (K1: conv0+GroupNorm+GELU, the bandwidth-bound stage)",
                    "achieved": round(gbs, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(gbs / hbm_peak, 4),
                    "peak_source": "measured" if peaks else "fallback", "launches_per_step": k1["launches"],
                    "algorithmic_bytes_per_launch": round(k1["bytes"] / max(1, k1["launches"])), "traffic": None}
        try:
            if args.workload == "c3":
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json")))
                hbm_roof["traffic"] = round(tj["kernels"]["conv0_gn_gelu"]["traffic_bytes_per_launch"])
                hbm_roof["traffic_source"] = "profiles/traffic_r02.json: " + tj["workload"]
        except Exception:
            pass
    if args.profile_json:
        with open(args.profile_json, "w") as f:
            json.dump({"kernels": ksum, "launch_list": [(k, fl, by, s.elapsed_time(e), d) for k, fl, by, d, s, e in prof.rec]}, f)

    # ---- CPU baseline + in-bench parity: the UNMODIFIED reference (fp32, host cores) on one whole batch per length decile;
    # the B200 memories of the SAME host batches (every row) are compared with the reference's
    cpu, parity = None, None
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(cores)
        fn, kind = reference_forward(M)
        picks, desc = decile_sample(batches, args.workload)
        t_cpu, a_cpu, checks = 0.0, 0.0, []
        with torch.no_grad():
            w0, l0 = host_batch([16000, 12000], seed=1)
            fn(w0, l0)                                               # warm-up (thread pool, allocator)
            for i, rows in picks:
                w, l = host[i]
                w, l = w[:rows], l[:rows]
                t = time.perf_counter()
                ref = fn(w, l)
                t_cpu += time.perf_counter() - t
                a_cpu += sum(batches[i][:rows]) / SR
                if rows == len(batches[i]):
                    got = enc(w.cuda(), l.cuda()).encoder_out.float().cpu()      # [M, B, 512]
                    d = (got.double() - ref.double())
                    per_row = (d.pow(2).sum((0, 2)).sqrt() / ref.double().pow(2).sum((0, 2)).sqrt())
                    checks.append({"batch": "%dx%d" % (len(batches[i]), max(batches[i])),
                                   "ragged_min_len": min(batches[i]),
                                   "rel_l2": round(float(d.norm() / ref.double().norm()), 6),
                                   "worst_row_rel_l2": round(float(per_row.max()), 6)})
        cpu = {"value": round(a_cpu / t_cpu, 3), "unit": "audio-s/s", "cores": cores, "kind": kind,
               "sample": "%s (%.0f audio-s, one pass), fp32, torch CPU %d threads" % (desc, a_cpu, cores)}
        if checks:
            tol = 1e-2 if dtype == torch.bfloat16 else 1e-5
            worst = max(c["worst_row_rel_l2"] for c in checks)
            parity = {"against": kind, "dtype": args.dtype, "tolerance_rel_l2": tol, "rows_checked": "every row of every batch",
                      "rel_l2": max(c["rel_l2"] for c in checks), "worst_row_rel_l2": worst, "ok": bool(worst <= tol),
                      "batches": checks}

    decode = None
    if args.decode:
        # encode + greedy decode of the first batch (reference: SequenceGenerator beam 1, max_len_a 0, max_len_b 200);
        # random-init hypotheses run into the forced EOS, i.e. the full 201 steps
        from chimera_st_b200.decoder import B200GreedyDecoder, B200BeamDecoder
        dsd = synth.make_decoder_state_dict(seed=1)
        dec = (B200GreedyDecoder(dsd, dtype=dtype, device="cuda") if args.beam == 1
               else B200BeamDecoder(dsd, beam=args.beam, dtype=dtype, device="cuda"))
        w0, l0 = dev[0]
        for _ in range(2):
            dec.generate(enc(w0, l0).encoder_out, max_len=200)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        mem0 = enc(w0, l0).encoder_out
        ev[1].record()
        hyp = dec.generate(mem0, max_len=200)
        ev[2].record()
        torch.cuda.synchronize()
        t_enc, t_dec = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        a0 = sum(batches[0]) / SR
        decode = {"batch": "%d utterances x %.1f s, M=%d" % (len(batches[0]), max(batches[0]) / SR, M),
                  "encode_ms": round(t_enc, 3), "decode_ms": round(t_dec, 3), "steps": dec.last_steps,
                  "us_per_step": round(1e3 * t_dec / max(1, dec.last_steps), 1),
                  "gpu_launches": dec.last_launches if args.beam == 1 else None,
                  "beam": args.beam,
                  "tokens": sum(len(h["tokens"]) for h in hyp) if args.beam == 1 else sum(len(hs[0]["tokens"]) for hs in hyp),
                  "encode_decode_audio_s_per_s": round(a0 / ((t_enc + t_dec) * 1e-3), 1),
                  "encode_only_audio_s_per_s": round(a0 / (t_enc * 1e-3), 1)}

    value = total_audio * args.steps / t_res
    h2d = sum(w.numel() * 4 + l.numel() * 8 for w, l in host)
    d2h = sum(o.numel() * 4 for o in out_host)
    line = {"metric": "encoded audio-sec/sec", "value": round(value, 1), "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(1e3 * t_res / args.steps, 3),
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "bf16" if dtype == torch.bfloat16 else "f32", "data": "synthetic", "config": cfg,
            "e2e": {"value": round(total_audio * args.steps / t_e2e, 1), "unit": "audio-s/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "wall_s": round(t_e2e_wall, 3),
                    "int16_wire": {"value": round(total_audio * args.steps / t_e2e16, 1),
                                   "h2d_bytes_per_step": sum(w.numel() * 2 + l.numel() * 8 for w, l in host)}},
            "gpu_launches": launches, "cuda_graph": not args.no_graph, "stream_lanes": lanes,
            "super_batches": {"count": len(supers), "frame_rows_target": enc.SUPER_ROWS if super_rows is None else super_rows,
                              "reference_batches_per_super_batch": round(len(dev) / max(1, len(supers)), 2)},
            "clocks": clocks,
            "roofline": roofline, "roofline_hbm_kernel": hbm_roof, "cpu_baseline": cpu, "parity": parity}
    if decode is not None:
        line["decode"] = decode
    if extras is not None:
        line["secondary"] = extras
    print(json.dumps(line))
    D.finalize()


if __name__ == "__main__":
    main()
