"""Greedy target-token decoding from the M memories on the B200 (BASELINE configs[3]).

Host-side mirror of the reference's generation stack for `beam_size = 1`:

    SequenceGenerator(models, tgt_dict, beam_size=1, max_len_a, max_len_b, min_len).generate(models, sample)
        fairseq/sequence_generator.py:17-540      -> B200GreedyGenerator.generate(models, sample)
    TransformerDecoder (incremental)  fairseq/models/transformer.py:640-838     -> B200GreedyDecoder
    TransformerDecoderLayer           fairseq/modules/transformer_layer.py:158-412

`B200GreedyDecoder` takes the reference's `decoder.*` state dict (same keys and shapes: q/k/v/out projections,
three LayerNorms and fc1/fc2 per layer, tied `embed_tokens` / `output_projection`), and produces, for every
utterance, the reference's hypothesis: token IDs ending in EOS, per-token log-probabilities and the
length-normalised score.  All arithmetic runs in the CUDA kernels behind `cst_dec_*` (csrc/decoder.cu); a decoding
step is 51 launches captured once into a CUDA graph and replayed -- the step index lives on the device, and the
host only polls the number of finished hypotheses every few steps.  No CPU fallback: a non-CUDA memory tensor raises
(the `lib` argument exists so that tests can drive the launch sequence against a host emulator of the C ABI).

Not reproduced: beam_size > 1, sampling, prefix tokens, n-gram blocking, temperature, LM fusion, alignment output.
"""
import ctypes as C
import math
import os

import torch

from . import _lib as L

PAD, EOS = 1, 2          # Dictionary defaults (fairseq/data/dictionary.py:26-31): bos 0, pad 1, eos 2, unk 3
DIM, FFN, HEADS, LAYERS = 512, 2048, 8, 6


def sinusoidal_positions(n_steps, dim=DIM, padding_idx=PAD):
    """Rows t = 0..n_steps-1 of SinusoidalPositionalEmbedding.get_embedding (sinusoidal_positional_embedding.py:38-59)
    at positions padding_idx + 1 + t (incremental decoding: `pos = padding_idx + seq_len`, :80-88)."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float) * -e)
    pos = torch.arange(padding_idx + 1, padding_idx + 1 + n_steps, dtype=torch.float)
    e = pos.unsqueeze(1) * e.unsqueeze(0)
    return torch.cat([torch.sin(e), torch.cos(e)], dim=1).contiguous()


def prepare_decoder_weights(sd, device, dtype, prefix="decoder."):
    """Reference `decoder.*` state dict -> fused device tensors.  Matrices in `dtype` (fp32 / bf16), biases and LayerNorm
    parameters fp32.  Only exact folds: the 1/8 query scaling (a power of two) into W_q, b_q; q|k|v concatenation."""
    if dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("decoder weight dtype must be float32 or bfloat16")

    def g(name):
        return sd[prefix + name].detach().float()

    def w(t):
        return t.to(device=device, dtype=dtype).contiguous()

    def f(t):
        return t.to(device=device, dtype=torch.float32).contiguous()

    # only the decoder variant the Chimera recipes train (pre-LN layers, sinusoidal positions, sqrt(d)-scaled tied
    # embeddings, no layernorm_embedding / project_in / project_out): anything else in the checkpoint is refused here so
    # that callers (the fairseq plugin) keep the reference SequenceGenerator instead of decoding wrong tokens silently
    import re
    known = re.compile(r"^(embed_tokens\.weight|output_projection\.weight|embed_positions\._float_tensor|version|"
                       r"layer_norm\.(weight|bias)|layers\.\d+\.((self_attn|encoder_attn)\.(q|k|v|out)_proj\.(weight|bias)|"
                       r"(self_attn_layer_norm|encoder_attn_layer_norm|final_layer_norm)\.(weight|bias)|fc[12]\.(weight|bias)))$")
    extra = [k[len(prefix):] for k in sd if k.startswith(prefix) and not known.match(k[len(prefix):])]
    if extra:
        raise NotImplementedError("unsupported decoder variant: unexpected checkpoint keys %s" % extra[:4])
    if prefix + "layer_norm.weight" not in sd:
        raise NotImplementedError("unsupported decoder variant: no final layer_norm (decoder_normalize_before=False)")
    P = {"layers": []}
    E = g("output_projection.weight") if prefix + "output_projection.weight" in sd else g("embed_tokens.weight")
    if not torch.equal(E, g("embed_tokens.weight")):
        raise NotImplementedError("untied decoder input/output embeddings")     # share_decoder_input_output_embed
    P["embed"] = w(E)
    P["vocab"] = E.shape[0]
    P["ln_g"], P["ln_b"] = f(g("layer_norm.weight")), f(g("layer_norm.bias"))
    n = 0
    while prefix + "layers.%d.fc1.weight" % n in sd:
        n += 1
    scale = (DIM // HEADS) ** -0.5
    for i in range(n):
        p = "layers.%d." % i
        sa, ea = p + "self_attn.", p + "encoder_attn."
        lay = {
            "ln1_g": f(g(p + "self_attn_layer_norm.weight")), "ln1_b": f(g(p + "self_attn_layer_norm.bias")),
            "qkv_w": w(torch.cat([g(sa + "q_proj.weight") * scale, g(sa + "k_proj.weight"), g(sa + "v_proj.weight")], 0)),
            "qkv_b": f(torch.cat([g(sa + "q_proj.bias") * scale, g(sa + "k_proj.bias"), g(sa + "v_proj.bias")], 0)),
            "so_w": w(g(sa + "out_proj.weight")), "so_b": f(g(sa + "out_proj.bias")),
            "ln2_g": f(g(p + "encoder_attn_layer_norm.weight")), "ln2_b": f(g(p + "encoder_attn_layer_norm.bias")),
            "xq_w": w(g(ea + "q_proj.weight") * scale), "xq_b": f(g(ea + "q_proj.bias") * scale),
            "xkv_w": w(torch.cat([g(ea + "k_proj.weight"), g(ea + "v_proj.weight")], 0)),
            "xkv_b": f(torch.cat([g(ea + "k_proj.bias"), g(ea + "v_proj.bias")], 0)),
            "xo_w": w(g(ea + "out_proj.weight")), "xo_b": f(g(ea + "out_proj.bias")),
            "ln3_g": f(g(p + "final_layer_norm.weight")), "ln3_b": f(g(p + "final_layer_norm.bias")),
            "fc1_w": w(g(p + "fc1.weight")), "fc1_b": f(g(p + "fc1.bias")),
            "fc2_w": w(g(p + "fc2.weight")), "fc2_b": f(g(p + "fc2.bias")),
        }
        P["layers"].append(lay)
    return P


class _DecodePlan:
    """Buffers + launch sequence (+ CUDA graph of one step) for a fixed (B hypotheses, M memories, max_len)."""

    def __init__(self, P, B, M, max_len, mem_dtype, device, lib, use_graph):
        self.P, self.B, self.M, self.max_len, self.lib = P, B, M, max_len, lib
        self.device = device
        self.use_graph = use_graph and device.type == "cuda"
        self.T = max_len + 2                                  # cache rows / token columns (EOS start + max_len + 1 picks)
        nl = len(P["layers"])
        z = dict(device=device, dtype=torch.float32)
        self.mem = torch.zeros(M * B, DIM, device=device, dtype=mem_dtype)       # row m*B + b  (= encoder_out [M,B,C])
        # K / V of the memories (per layer) and the self-attention cache: bf16 in the 16-bit mode (half the HBM bytes of
        # the only tensors a step streams from DRAM), fp32 in the fp32 parity mode
        kv = dict(device=device, dtype=P["embed"].dtype)
        self.xk = torch.zeros(nl, M * B, DIM, **kv)
        self.xv = torch.zeros(nl, M * B, DIM, **kv)
        self.kc = torch.zeros(nl, B, self.T, DIM, **kv)
        self.vc = torch.zeros(nl, B, self.T, DIM, **kv)
        self.x = torch.zeros(B, DIM, **z)
        self.q = torch.zeros(B, DIM, **z)
        self.a = torch.zeros(B, DIM, **z)
        self.h = torch.zeros(B, FFN, **z)
        self.logits = torch.zeros(B, P["vocab"], **z)
        self.pos = sinusoidal_positions(self.T).to(device)
        self.tokens = torch.zeros(B, self.T, device=device, dtype=torch.int32)
        self.pos_scores = torch.zeros(B, self.T, **z)
        self.done = torch.zeros(B, device=device, dtype=torch.int32)
        self.out_len = torch.zeros(B, device=device, dtype=torch.int32)
        self.counters = torch.zeros(4, device=device, dtype=torch.int32)
        self.graph = None
        self.graph_chunk = None                               # CHUNK steps in one graph (run_async)
        self.min_len = 1
        self.launches_per_step = 0
        self.total_launches = 0
        self.n_begin = 0

    # ---- launches -------------------------------------------------------------------------------------
    def _st(self):
        return L.stream_ptr() if self.device.type == "cuda" else 0

    def _linear(self, A, W, bias, outs, M, N, K, ln=None, residual=None, act=L.ACT_NONE, ldo=None, step_stride=None,
                use_step=False):
        p = L.DecLinearParams()
        p.A, p.W, p.bias = A.data_ptr(), W.data_ptr(), L.ptr(bias)
        p.ln_gamma, p.ln_beta = (ln[0].data_ptr(), ln[1].data_ptr()) if ln is not None else (0, 0)
        p.residual = L.ptr(residual)
        p.lda, p.ldr = K, N
        n_seg = len(outs)
        for s in range(3):
            p.out[s] = outs[s].data_ptr() if s < n_seg else 0
            p.out_dtype[s] = L.DT[outs[s].dtype] if s < n_seg else L.F32
            p.ldo[s] = (ldo[s] if ldo is not None else N // n_seg) if s < n_seg else 0
            p.step_stride[s] = step_stride[s] if (step_stride is not None and s < n_seg) else 0
        p.step = self.counters.data_ptr() if use_step else 0
        p.a_dtype, p.w_dtype = L.DT[A.dtype], L.DT[W.dtype]
        p.M, p.N, p.K, p.n_seg, p.act = M, N, K, n_seg, act
        L.check(self.lib.cst_dec_linear(C.byref(p), self._st()))
        self.n_launch += 1

    def _attention(self, k, v, kv_bs, kv_rs, n_keys, n_max, use_step):
        L.check(self.lib.cst_dec_attention(self.q.data_ptr(), DIM, k.data_ptr(), v.data_ptr(), L.DT[k.dtype], kv_bs, kv_rs,
                                           self.a.data_ptr(), DIM, self.B, HEADS, n_keys, n_max,
                                           self.counters.data_ptr() if use_step else 0, self._st()))
        self.n_launch += 1

    def begin(self, memories):
        """memories [M,B,512] -> cross-attention K/V of every layer (static_kv, multihead_attention.py:213-221);
        reset tokens / counters."""
        B, M, P = self.B, self.M, self.P
        self.n_launch = 0
        self.mem.copy_(memories.reshape(M * B, DIM))
        self.tokens.fill_(EOS)                               # tokens[:, 0] = EOS (sequence_generator.py:247-252)
        self.pos_scores.zero_()
        self.done.zero_()
        self.out_len.zero_()
        self.counters.zero_()
        for i, lay in enumerate(P["layers"]):
            self._linear(self.mem, lay["xkv_w"], lay["xkv_b"], [self.xk[i], self.xv[i]], M * B, 2 * DIM, DIM)

    def _step(self):
        B, M, P, T = self.B, self.M, self.P, self.T
        self.n_launch = 0
        L.check(self.lib.cst_dec_embed(self.tokens.data_ptr(), T, P["embed"].data_ptr(), L.DT[P["embed"].dtype],
                                       self.pos.data_ptr(), math.sqrt(DIM), self.x.data_ptr(), B, DIM,
                                       self.counters.data_ptr(), self._st()))
        self.n_launch += 1
        for i, lay in enumerate(P["layers"]):
            # self-attention block: x += Out(attn(LN1 x)); K/V rows of this step appended to the cache by the projection
            self._linear(self.x, lay["qkv_w"], lay["qkv_b"], [self.q, self.kc[i], self.vc[i]], B, 3 * DIM, DIM,
                         ln=(lay["ln1_g"], lay["ln1_b"]), ldo=[DIM, T * DIM, T * DIM], step_stride=[0, DIM, DIM],
                         use_step=True)
            self._attention(self.kc[i], self.vc[i], T * DIM, DIM, 0, T, True)
            self._linear(self.a, lay["so_w"], lay["so_b"], [self.x], B, DIM, DIM, residual=self.x)
            # encoder-decoder attention over the M memories
            self._linear(self.x, lay["xq_w"], lay["xq_b"], [self.q], B, DIM, DIM, ln=(lay["ln2_g"], lay["ln2_b"]))
            self._attention(self.xk[i], self.xv[i], DIM, B * DIM, M, M, False)
            self._linear(self.a, lay["xo_w"], lay["xo_b"], [self.x], B, DIM, DIM, residual=self.x)
            # feed-forward
            self._linear(self.x, lay["fc1_w"], lay["fc1_b"], [self.h], B, FFN, DIM, ln=(lay["ln3_g"], lay["ln3_b"]),
                         act=L.ACT_RELU)
            self._linear(self.h, lay["fc2_w"], lay["fc2_b"], [self.x], B, DIM, FFN, residual=self.x)
        self._linear(self.x, P["embed"], None, [self.logits], B, P["vocab"], DIM, ln=(P["ln_g"], P["ln_b"]))
        L.check(self.lib.cst_dec_select(self.logits.data_ptr(), P["vocab"], B, self.tokens.data_ptr(), T,
                                        self.pos_scores.data_ptr(), T, self.done.data_ptr(), self.out_len.data_ptr(),
                                        self.counters.data_ptr(), self.max_len, self.min_len, PAD, EOS, self._st()))
        self.n_launch += 1
        self.launches_per_step = self.n_launch

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step()

    def prepare(self, memories, min_len=1):
        """Reset the device state for a new batch (and capture the step graph on first use), on the current stream."""
        if self.graph is not None and min_len != self.min_len:
            self.graph = self.graph_chunk = None               # min_len is a kernel argument baked into the graph
        self.min_len = min_len
        self.begin(memories)
        if self.use_graph and self.graph is None:
            # warm-up step (module loading, function attributes), then capture; the step index and token state are
            # device state, so rewind them afterwards (capturing does not execute anything)
            self._step()
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step()
            self.graph = g
            self.begin(memories)
        self.n_begin = self.n_launch

    CHUNK = int(os.environ.get("CST_DEC_CHUNK", "8"))     # decoding steps per graph of run_async

    def run_async(self, memories, min_len=1):
        """Enqueue a WHOLE decode (max_len + 1 steps, rounded up to CHUNK; steps past the last one are no-ops of the select kernel)
        on the current stream without any host synchronisation: ceil((max_len + 1) / CHUNK) replays of a CHUNK-step graph.  Used to
        run the latency-bound decode of one batch underneath the encoder / the decodes of other batches (decoder.generate_async).
        -> number of steps enqueued."""
        self.prepare(memories, min_len)
        if self.use_graph and self.graph_chunk is None:
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(self.CHUNK):
                    self._step()
            self.graph_chunk = g
            self.begin(memories)
            self.n_begin = self.n_launch
        n = -(-(self.max_len + 1) // self.CHUNK)
        for _ in range(n):
            if self.graph_chunk is not None:
                self.graph_chunk.replay()
            else:
                for _ in range(self.CHUNK):
                    self._step()
        self.total_launches = self.n_begin + n * self.CHUNK * self.launches_per_step
        return n * self.CHUNK

    def run(self, memories, min_len=1, poll=8):
        """-> number of steps issued."""
        self.prepare(memories, min_len)
        steps = 0
        for s in range(self.max_len + 1):
            self.step()
            steps += 1
            if s == self.max_len or (s % poll) == poll - 1:
                if int(self.counters[2].item()) >= self.B:     # the only host synchronisation of the loop
                    break
        self.total_launches = self.n_begin + steps * self.launches_per_step
        return steps


class B200GreedyDecoder:
    """decoder.* state dict + memories -> greedy hypotheses (see module docstring).

    Stream lanes (`n_lanes` > 1, off by default): hypotheses are independent, so `generate` can split the batch into
    sub-batches, each with its own buffers, CUDA graph and stream, whose dependent kernel chains interleave on the GPU;
    results are bit-identical to one lane (no kernel mixes rows).  Measured on B200 at the C4 shape (64 hypotheses,
    bf16): 1 / 2 / 4 / 8 lanes = 75 / 79 / 92 / 144 ms per batch -- one host thread pays ~100 us per 51-node graph
    launch, so with several lanes the host, not the GPU, paces the steps.  Kept as a lever for multi-threaded hosts."""

    def __init__(self, state_dict, dtype=torch.float32, device="cuda", prefix="decoder.", use_graph=True, lib=None,
                 n_lanes=None):
        self.device = torch.device(device)
        if lib is None:
            if self.device.type != "cuda":
                raise L.CstError("B200GreedyDecoder needs a CUDA device (there is no CPU fallback)")
            lib = L.load()
        self.lib = lib
        self.P = prepare_decoder_weights(state_dict, self.device, dtype, prefix)
        self.use_graph = use_graph
        self.n_lanes = n_lanes if n_lanes is not None else int(os.environ.get("CST_DEC_LANES", "1"))
        self._plans = {}
        self._streams = []
        self.last_steps = 0
        self.last_launches = 0
        self.last_lanes = 1

    def _plan(self, B, M, max_len, mem_dtype, lane=0):
        key = (B, M, max_len, mem_dtype, lane)
        if key not in self._plans:
            if len(self._plans) >= 16:
                self._plans.pop(next(iter(self._plans)))
            self._plans[key] = _DecodePlan(self.P, B, M, max_len, mem_dtype, self.device, self.lib, self.use_graph)
        return self._plans[key]

    def _lane_split(self, B, n_lanes):
        """Sub-batches of >= 8 hypotheses, sizes as equal as possible."""
        n = max(1, min(n_lanes, B // 8)) if self.device.type == "cuda" else 1
        base, extra = divmod(B, n)
        sizes = [base + (1 if i < extra else 0) for i in range(n)]
        offs = [sum(sizes[:i]) for i in range(n)]
        return list(zip(offs, sizes))

    @torch.no_grad()
    def generate(self, memories, max_len=200, min_len=1, n_lanes=None, poll=8):
        """memories: encoder_out [M,B,512] (fp32 or bf16, on the decoder's device).
        max_len = int(max_len_a * src_len + max_len_b) as in sequence_generator.py:222-230.
        -> list over utterances of {"tokens": LongTensor [n] ending in EOS, "score": float (sum of log-probs / n),
           "positional_scores": FloatTensor [n]} -- the reference's hypothesis dict (sequence_generator.py:636-648)."""
        if memories.device.type != self.device.type:
            raise L.CstError("memories must live on %s (no CPU fallback)" % self.device)
        if memories.dim() != 3 or memories.shape[2] != DIM:
            raise ValueError("memories must be [M, B, %d]" % DIM)
        if memories.dtype == torch.float16:          # fairseq --fp16 models hand over half memories; the kernels take fp32 / bf16
            memories = memories.float()
        if memories.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("memories must be float32, bfloat16 or float16")
        M, B = memories.shape[0], memories.shape[1]
        max_len = int(max_len)
        split = self._lane_split(B, self.n_lanes if n_lanes is None else n_lanes)
        self.last_lanes = len(split)
        if len(split) == 1:
            plan = self._plan(B, M, max_len, memories.dtype)
            self.last_steps = plan.run(memories.contiguous(), min_len=min_len, poll=poll)
            plans = [plan]
        else:
            plans = self._run_lanes(memories, split, M, max_len, min_len, poll)
        self.last_launches = sum(p.total_launches for p in plans)
        return self._hypotheses(plans)

    @staticmethod
    def _hypotheses(plans):
        out = []
        for plan in plans:
            toks, lens, ps = plan.tokens.cpu(), plan.out_len.cpu(), plan.pos_scores.cpu()
            for b in range(plan.B):
                n = int(lens[b])
                sc = ps[b, :n].clone()
                out.append({"tokens": toks[b, 1:n + 1].long(), "score": float(sc.sum() / max(n, 1)), "attention": None,
                            "alignment": torch.empty(0), "positional_scores": sc})
        return out

    @torch.no_grad()
    def generate_async(self, memories, max_len=200, min_len=1, lane=0):
        """Throughput form: enqueue the whole greedy decode of `memories` on lane `lane`'s own stream and return at once (no host
        synchronisation; `collect(handle)` waits and returns the hypotheses of `generate`).  A decoding step is a latency chain of 51
        short kernels that leaves most SMs idle, so the decodes of several batches -- and the encoder pass of the next batch on the
        caller's stream -- overlap almost for free; one lane holds one batch at a time (collect before reusing it)."""
        if memories.device.type != self.device.type or self.device.type != "cuda":
            raise L.CstError("generate_async needs CUDA memories (no CPU fallback)")
        if memories.dtype == torch.float16:
            memories = memories.float()
        M, B = memories.shape[0], memories.shape[1]
        while len(self._streams) <= lane:
            self._streams.append(torch.cuda.Stream(device=self.device))
        st = self._streams[lane]
        plan = self._plan(B, M, int(max_len), memories.dtype, lane=("async", lane))
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            steps = plan.run_async(memories.contiguous(), min_len=min_len)
        memories.record_stream(st)
        return plan, st, steps

    def collect(self, handle):
        plan, st, steps = handle
        st.synchronize()
        self.last_steps, self.last_launches, self.last_lanes = steps, plan.total_launches, 1
        return self._hypotheses([plan])

    def _run_lanes(self, memories, split, M, max_len, min_len, poll):
        while len(self._streams) < len(split):
            self._streams.append(torch.cuda.Stream(device=self.device))
        cur = torch.cuda.current_stream()
        plans = [self._plan(nb, M, max_len, memories.dtype, lane=i) for i, (_, nb) in enumerate(split)]
        for i, (plan, (b0, nb)) in enumerate(zip(plans, split)):
            st = self._streams[i]
            st.wait_stream(cur)                                   # the memories were produced on the caller's stream
            with torch.cuda.stream(st):
                plan.prepare(memories[:, b0:b0 + nb].reshape(M * nb, DIM), min_len)
        steps = 0
        for s in range(max_len + 1):
            for i, plan in enumerate(plans):
                with torch.cuda.stream(self._streams[i]):
                    plan.step()
            steps += 1
            if s == max_len or (s % poll) == poll - 1:
                fin = 0
                for i, plan in enumerate(plans):
                    with torch.cuda.stream(self._streams[i]):
                        fin += int(plan.counters[2].item())
                if fin >= sum(nb for _, nb in split):
                    break
        for i, plan in enumerate(plans):
            plan.total_launches = plan.n_begin + steps * plan.launches_per_step
            cur.wait_stream(self._streams[i])
        self.last_steps = steps
        return plans


class _BeamPlan(_DecodePlan):
    """Buffers + launch sequence of beam search for B sentences x K beams (R = B*K decoder rows).  Tokens, cumulative scores
    and the cache-history table are double-buffered (the select kernel permutes rows from one set into the other), so the
    replayed CUDA graph holds TWO steps (parity 0 then 1)."""

    def __init__(self, P, B, K, M, max_len, mem_dtype, device, lib, use_graph):
        super().__init__(P, B * K, M, max_len, mem_dtype, device, lib, use_graph)
        self.nsent, self.K, self.R = B, K, B * K
        R, T = self.R, self.T
        # the K beams of a sentence attend the same memories: K / V are projected and kept once per SENTENCE and the memory
        # attention maps decoder row r to set r / K (cst_dec_attention_grouped) -- 1/K of the memory, no repeat_interleave copy
        nl = len(P["layers"])
        self.mem = torch.zeros(M * B, DIM, device=device, dtype=mem_dtype)
        self.xk = torch.zeros(nl, M * B, DIM, device=device, dtype=P["embed"].dtype)
        self.xv = torch.zeros(nl, M * B, DIM, device=device, dtype=P["embed"].dtype)
        zi = dict(device=device, dtype=torch.int32)
        zf = dict(device=device, dtype=torch.float32)
        self.tok = [torch.zeros(R, T, **zi) for _ in range(2)]
        self.sc = [torch.zeros(R, T, **zf) for _ in range(2)]
        self.hist = [torch.zeros(R, T, **zi) for _ in range(2)]
        self.ignore = torch.zeros(R, **zi)
        self.fin_tokens = torch.zeros(B, K, T, **zi)
        self.fin_pos = torch.zeros(B, K, T, **zf)
        self.fin_score = torch.zeros(B, K, **zf)
        self.fin_len = torch.zeros(B, K, **zi)
        self.n_final = torch.zeros(B, **zi)
        self.finished = torch.zeros(B, **zi)
        self.len_penalty = 1.0

    def begin_beam(self, memories):
        """memories [M, B, 512]: every sentence's memories are repeated for its K beams (reorder_encoder_out with
        new_order = arange(B).repeat_interleave(K), sequence_generator.py:239-243)."""
        M, B, P = self.M, self.nsent, self.P
        self.n_launch = 0
        self.mem.copy_(memories.reshape(M * B, DIM))
        for t in (self.tokens, self.pos_scores, self.done, self.out_len, self.counters):
            t.zero_()
        for i, lay in enumerate(P["layers"]):
            self._linear(self.mem, lay["xkv_w"], lay["xkv_b"], [self.xk[i], self.xv[i]], M * B, 2 * DIM, DIM)
        for t in self.tok:
            t.fill_(PAD)
            t[:, 0] = EOS
        for t in self.sc + self.hist + [self.ignore, self.fin_tokens, self.fin_pos, self.fin_score, self.fin_len, self.n_final,
                                        self.finished]:
            t.zero_()

    def _step_beam(self, parity):
        R, M, P, T = self.R, self.M, self.P, self.T
        tok_in, tok_out = self.tok[parity], self.tok[1 - parity]
        n0 = self.n_launch
        L.check(self.lib.cst_dec_embed(tok_in.data_ptr(), T, P["embed"].data_ptr(), L.DT[P["embed"].dtype],
                                       self.pos.data_ptr(), math.sqrt(DIM), self.x.data_ptr(), R, DIM,
                                       self.counters.data_ptr(), self._st()))
        self.n_launch += 1
        for i, lay in enumerate(P["layers"]):
            self._linear(self.x, lay["qkv_w"], lay["qkv_b"], [self.q, self.kc[i], self.vc[i]], R, 3 * DIM, DIM,
                         ln=(lay["ln1_g"], lay["ln1_b"]), ldo=[DIM, T * DIM, T * DIM], step_stride=[0, DIM, DIM],
                         use_step=True)
            L.check(self.lib.cst_dec_attention_beam(self.q.data_ptr(), DIM, self.kc[i].data_ptr(), self.vc[i].data_ptr(),
                                                    L.DT[self.kc.dtype], T * DIM, DIM, self.a.data_ptr(), DIM, R, HEADS, T,
                                                    self.hist[parity].data_ptr(), T, self.counters.data_ptr(), self._st()))
            self.n_launch += 1
            self._linear(self.a, lay["so_w"], lay["so_b"], [self.x], R, DIM, DIM, residual=self.x)
            self._linear(self.x, lay["xq_w"], lay["xq_b"], [self.q], R, DIM, DIM, ln=(lay["ln2_g"], lay["ln2_b"]))
            L.check(self.lib.cst_dec_attention_grouped(self.q.data_ptr(), DIM, self.xk[i].data_ptr(), self.xv[i].data_ptr(),
                                                       L.DT[self.xk.dtype], DIM, self.nsent * DIM, self.a.data_ptr(), DIM, R, HEADS, M, M,
                                                       0, self.K, self._st()))
            self.n_launch += 1
            self._linear(self.a, lay["xo_w"], lay["xo_b"], [self.x], R, DIM, DIM, residual=self.x)
            self._linear(self.x, lay["fc1_w"], lay["fc1_b"], [self.h], R, FFN, DIM, ln=(lay["ln3_g"], lay["ln3_b"]),
                         act=L.ACT_RELU)
            self._linear(self.h, lay["fc2_w"], lay["fc2_b"], [self.x], R, DIM, FFN, residual=self.x)
        self._linear(self.x, P["embed"], None, [self.logits], R, P["vocab"], DIM, ln=(P["ln_g"], P["ln_b"]))
        p = L.DecBeamParams()
        p.logits = self.logits.data_ptr()
        p.tok_in, p.tok_out = tok_in.data_ptr(), tok_out.data_ptr()
        p.sc_in, p.sc_out = self.sc[parity].data_ptr(), self.sc[1 - parity].data_ptr()
        p.hist_in, p.hist_out = self.hist[parity].data_ptr(), self.hist[1 - parity].data_ptr()
        p.ignore = self.ignore.data_ptr()
        p.fin_tokens, p.fin_pos, p.fin_score = self.fin_tokens.data_ptr(), self.fin_pos.data_ptr(), self.fin_score.data_ptr()
        p.fin_len, p.n_final, p.finished = self.fin_len.data_ptr(), self.n_final.data_ptr(), self.finished.data_ptr()
        p.counters = self.counters.data_ptr()
        p.B, p.K, p.V, p.T = self.nsent, self.K, P["vocab"], T
        p.max_len, p.min_len, p.pad, p.eos, p.len_penalty = self.max_len, self.min_len, PAD, EOS, self.len_penalty
        L.check(self.lib.cst_dec_beam_select(C.byref(p), self._st()))
        self.n_launch += 1
        self.launches_per_step = self.n_launch - n0

    def run_beam(self, memories, min_len=1, poll=4):
        if self.graph is not None and min_len != self.min_len:
            self.graph = None
        self.min_len = min_len
        self.begin_beam(memories)
        if self.use_graph and self.graph is None:
            self._step_beam(0)
            self._step_beam(1)
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_beam(0)
                self._step_beam(1)
            self.graph = g
            self.begin_beam(memories)
        steps = 0
        while steps <= self.max_len:
            if self.graph is not None:
                self.graph.replay()                            # two steps; a step past max_len is a no-op in the select kernel
                steps += 2
            else:
                self._step_beam(steps & 1)
                steps += 1
            if steps > self.max_len or (steps // 2) % poll == 0:
                if int(self.counters[2].item()) >= self.nsent:
                    break
        return steps


class B200BeamDecoder(B200GreedyDecoder):
    """Beam search (SequenceGenerator with beam_size K <= 8, default options) on the cst_dec_* kernels plus
    `cst_dec_attention_beam` / `cst_dec_beam_select`.  The K/V cache is never re-ordered: a per-row history table says
    which physical cache row holds each past position of a beam.  Checked against the reference generator's beam-5
    hypotheses (tests/golden/beam.npz) on the ABI emulator (CPU) and on the B200 (tests/test_gpu_beam.py)."""

    MAX_BEAM_PLANS = 4

    def __init__(self, state_dict, beam=5, **kw):
        super().__init__(state_dict, **kw)
        if not 1 <= beam <= 8:
            raise ValueError("beam must be in 1..8")
        self.beam = beam

    @torch.no_grad()
    def generate(self, memories, max_len=200, min_len=1, len_penalty=1.0):
        """-> per sentence a list (best first) of hypothesis dicts, as SequenceGenerator.generate returns them."""
        if memories.device.type != self.device.type:
            raise L.CstError("memories must live on %s (no CPU fallback)" % self.device)
        if memories.dtype == torch.float16:          # fairseq --fp16 models hand over half memories; the kernels take fp32 / bf16
            memories = memories.float()
        M, B = memories.shape[0], memories.shape[1]
        key = ("beam", B, self.beam, M, int(max_len), memories.dtype)
        if key not in self._plans:
            # bounded LRU: a beam plan owns 2*6*(B*K)*(max_len+2)*512 cache elements + a CUDA graph, and fairseq-generate
            # presents a new (B, max_len) for most batches of a real test set
            while len(self._plans) >= self.MAX_BEAM_PLANS:
                self._plans.pop(next(iter(self._plans)))
            self._plans[key] = _BeamPlan(self.P, B, self.beam, M, int(max_len), memories.dtype, self.device, self.lib, self.use_graph)
        plan = self._plans[key] = self._plans.pop(key)              # most recently used last
        if plan.graph is not None and float(len_penalty) != plan.len_penalty:
            plan.graph = None          # len_penalty is a by-value argument of cst_dec_beam_select: baked into the captured graph
        plan.len_penalty = float(len_penalty)
        self.last_steps = plan.run_beam(memories.contiguous(), min_len=min_len)
        ft, fp, fs, fl, nf = (t.cpu() for t in (plan.fin_tokens, plan.fin_pos, plan.fin_score, plan.fin_len, plan.n_final))
        out = []
        for b in range(B):
            n = int(nf[b])
            order = torch.sort(fs[b, :n], descending=True).indices.tolist()      # "sort by score descending", :533-540
            out.append([{"tokens": ft[b, k, :int(fl[b, k])].long(), "score": float(fs[b, k]), "attention": None,
                         "alignment": torch.empty(0), "positional_scores": fp[b, k, :int(fl[b, k])].clone()} for k in order])
        return out


class B200GreedyGenerator:
    """Drop-in for `SequenceGenerator(models, tgt_dict, beam_size<=8, ...)` (fairseq/sequence_generator.py:17-100):
    `generate(models, sample)` runs the model's own encoder (the B200 encoder when the plugin is active) and decodes on
    the GPU -- greedily for beam_size 1, with `B200BeamDecoder` for 2..8.  Returns the reference structure: per sentence a
    list (best first) of hypothesis dicts.  Options that change the search beyond that (sampling, unk penalty, temperature,
    n-gram blocking, prefixes, unnormalised scores ...) raise."""

    def __init__(self, models, tgt_dict=None, beam_size=1, max_len_a=0.0, max_len_b=200, min_len=1, normalize_scores=True,
                 len_penalty=1.0, unk_penalty=0.0, temperature=1.0, match_source_len=False, no_repeat_ngram_size=0,
                 symbols_to_strip_from_output=None, dtype=None, lib=None):
        models = list(models) if isinstance(models, (list, tuple)) else [models]
        if len(models) != 1:
            raise NotImplementedError("model ensembles")
        if (not 1 <= beam_size <= 8 or not normalize_scores or unk_penalty != 0 or temperature != 1.0
                or match_source_len or no_repeat_ngram_size):
            raise NotImplementedError("only plain beam search (beam <= 8, default options) runs on the B200 decoder")
        self.model = models[0]
        margs = getattr(self.model, "args", None)
        if margs is not None:
            bad = [k for k, want in (("decoder_learned_pos", False), ("no_scale_embedding", False), ("layernorm_embedding", False),
                                     ("decoder_normalize_before", True), ("no_token_positional_embeddings", False),
                                     ("decoder_embed_dim", DIM), ("decoder_ffn_embed_dim", FFN), ("decoder_attention_heads", HEADS),
                                     ("adaptive_softmax_cutoff", None), ("cross_self_attention", False))
                   if getattr(margs, k, want) != want]
            if bad:
                raise NotImplementedError("decoder options not supported by the B200 decoder: %s" % bad)
        self.pad, self.eos = PAD, EOS
        if tgt_dict is not None and (tgt_dict.pad(), tgt_dict.eos()) != (PAD, EOS):
            raise NotImplementedError("non-default pad/eos indices")
        self.symbols_to_strip_from_output = (set(symbols_to_strip_from_output) | {EOS}
                                             if symbols_to_strip_from_output is not None else {EOS})
        self.beam_size, self.max_len_a, self.max_len_b, self.min_len = beam_size, max_len_a, max_len_b, min_len
        self.len_penalty = float(len_penalty)
        dec_sd = {k: v for k, v in self.model.state_dict().items() if k.startswith("decoder.")}
        p = next(self.model.decoder.parameters())
        half = p.dtype in (torch.bfloat16, torch.float16)
        wdt = dtype or (torch.bfloat16 if half else torch.float32)
        if beam_size == 1:
            self.decoder = B200GreedyDecoder(dec_sd, dtype=wdt, device=p.device, lib=lib)
        else:
            self.decoder = B200BeamDecoder(dec_sd, beam=beam_size, dtype=wdt, device=p.device, lib=lib)

    def cuda(self):
        return self

    @torch.no_grad()
    def generate(self, models, sample, prefix_tokens=None, constraints=None, bos_token=None, **kwargs):
        if prefix_tokens is not None or constraints is not None or bos_token is not None:
            raise NotImplementedError("prefix tokens / constraints / bos override")
        net_input = sample["net_input"]
        enc = self.model.encoder(**{k: v for k, v in net_input.items() if k != "prev_output_tokens"})
        src_len = net_input["src_tokens"].shape[1]
        max_len = min(int(self.max_len_a * src_len + self.max_len_b), self.model.max_decoder_positions() - 1)
        if self.beam_size > 1:
            return self.decoder.generate(enc.encoder_out, max_len=max_len, min_len=self.min_len, len_penalty=self.len_penalty)
        hyps = self.decoder.generate(enc.encoder_out, max_len=max_len, min_len=self.min_len)
        for h in hyps:                                             # greedy: the normalised score with the requested length penalty
            n = len(h["tokens"])
            h["score"] = float(h["positional_scores"].sum()) / max(n, 1) ** self.len_penalty
        return [[h] for h in hyps]


B200Generator = B200GreedyGenerator        # the class serves beam widths 1..8; the first name is kept for callers
