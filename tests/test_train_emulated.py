"""Host logic of the training step (BASELINE configs[4]) without a GPU: the REAL `EncoderTrainStep` / `TextTrainPass` / `FusedAdam` launch
sequences run on CPU tensors against the host emulator of the C ABI (tests/emu.py), and every gradient is compared with torch autograd
through the oracle restatement.  This pins what lives above the kernels -- frame geometry, tape layout, operand transposes and the
split-K chunk layout, row remaps around the subsampler and the positional convolution, the mapping from kernel-layout gradients back to
the reference's parameter names, gradient accumulation over the audio and text passes -- on the box that has no GPU; the kernels
themselves are checked against the same autograd reference in tests/test_gpu_backward.py."""
import os

import numpy as np
import torch

import chimera_st_b200  # noqa: F401
from chimera_st_b200 import synth, _lib as L
from chimera_st_b200.train import EncoderTrainStep, TextTrainPass, FusedAdam
from oracle import chimera_oracle as O, adam_oracle
from conftest import GOLDEN, rel_l2
from emu import EmuLib, oracle_dropout_hook
import pytest


def _pinned(masks, fn):
    """Run `fn` with torch.relu pinned to the sign pattern OUR forward pass saw (pre-activations within rounding of zero legitimately
    flip between two fp32 evaluations; see tests/test_gpu_backward.py::_oracle_grads)."""
    calls, orig = [], torch.relu

    def relu(z):
        calls.append(1)
        return z * masks[len(calls) - 1].to(z.dtype)
    torch.relu = relu
    try:
        return fn()
    finally:
        torch.relu = orig


def _masks(T, B, rows_per_seg, n_rows, M):
    m = [t["z"][:B * rows_per_seg].view(B, rows_per_seg, -1)[:, :n_rows] > 0 for t in T["enc"]]
    return m + [t["z"].view(B, M, -1) > 0 for t in T["mem"]]


def _autograd(sd, run, R, masks):
    sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() else v) for k, v in sd.items()}
    mem = _pinned(masks, lambda: run(sdg))
    (mem * R).sum().backward()
    return mem.detach(), {k: v.grad for k, v in sdg.items() if v.is_floating_point() and v.grad is not None}


def _compare(G, ref, tol=2e-4):
    missing = [k for k in ref if k not in G and float(ref[k].abs().max()) > 0]
    assert not missing, missing
    bad = {}
    for k, v in G.items():
        if k.endswith("k_proj.bias"):                              # mathematically zero: rounding noise on both sides
            scale = float(ref[k.replace("k_proj", "q_proj")].abs().max())
            if not float(v.abs().max()) < 1e-4 * scale:
                bad[k] = float(v.abs().max())
            continue
        e = rel_l2(v.reshape(ref[k].shape), ref[k])
        if not e < tol:
            bad[k] = e
    assert not bad, bad


def test_training_step_requires_the_cuda_library_on_cpu():
    sd = synth.make_state_dict(seed=0, interlingua_length=4)
    with pytest.raises(L.CstError):
        EncoderTrainStep(sd, 1, 4000, device="cpu")
    with pytest.raises(L.CstError):
        FusedAdam({"w": torch.zeros(4)})


def test_emulated_training_step_matches_autograd_through_the_oracle():
    torch.set_num_threads(8)
    lens = [5200, 3900]
    sd = synth.make_state_dict(seed=0, interlingua_length=8, dead_heads=False)
    wave, tl = synth.make_waveforms(lens, seed=31)
    R = torch.randn(8, len(lens), 512, generator=torch.Generator().manual_seed(1))
    emu = EmuLib()
    step = EncoderTrainStep(sd, len(lens), wave.shape[1], device="cpu", feature_grad_mult=0.1, lib=emu)
    g = step.g
    mem, G = step.forward_backward(wave, tl, R)
    ref_mem, ref = _autograd(sd, lambda s: O.encoder_forward(s, wave, tl)[0], R, _masks(step.T, g.B, g.T2a, g.T2, 8))
    assert rel_l2(mem, ref_mem) < 1e-5
    # GradMultiply(0.1) on the feature extractor (wav2vec2.py:530-532): the oracle's autograd graph has no scaling, so undo ours
    undo = lambda G_: {k: (v * 10 if ".feature_extractor." in k else v) for k, v in G_.items()}     # noqa: E731
    _compare(undo(G), ref)
    # the launch sequence used every derivative entry point
    for name in ("transpose", "colsum", "act_bwd", "layernorm_bwd", "attention_bwd", "col2im", "rows_remap", "conv0_bwd"):
        assert name in emu.calls, name
    # LayerDrop (wav2vec2.py:835-838): dropped layers are the identity and their parameters get no gradient
    skip = frozenset({1, 7})
    mem2 = step.forward(wave, tl, skip_w2v_layers=skip)
    G2 = step.backward(R)
    orig_layer = O.w2v_layer
    O.w2v_layer = lambda sd_, i, x, m: x if i in skip else orig_layer(sd_, i, x, m)
    try:
        ref_mem2, ref2 = _autograd(sd, lambda s: O.encoder_forward(s, wave, tl)[0], R, _masks(step.T, g.B, g.T2a, g.T2, 8))
    finally:
        O.w2v_layer = orig_layer
    assert rel_l2(mem2, ref_mem2) < 1e-5
    assert not any(".encoder.layers.1." in k or ".encoder.layers.7." in k for k in G2)
    _compare(undo(G2), ref2)


def test_emulated_text_pass_accumulates_into_the_audio_gradients():
    torch.set_num_threads(8)
    V = 60
    sd = synth.make_state_dict(seed=2, interlingua_length=8, dead_heads=False, text_vocab=V)
    wave, tl = synth.make_waveforms([4100], seed=5)
    gen = torch.Generator().manual_seed(9)
    tok_len = torch.tensor([7, 4])
    tokens = torch.randint(4, V, (2, 7), generator=gen)
    tokens[1, 4:] = 1                                              # pad
    Ra = torch.randn(8, 1, 512, generator=gen)
    Rt = torch.randn(8, 2, 512, generator=gen)
    emu = EmuLib()
    step = EncoderTrainStep(sd, 1, wave.shape[1], device="cpu", feature_grad_mult=1.0, lib=emu)
    text = TextTrainPass(step, 2, 7)
    step.forward(wave, tl)
    mem_t = text.forward(tokens, tok_len)
    G = step.backward(Ra)
    G = text.backward(Rt, G)
    g = step.g
    masks = _masks(step.T, 1, g.T2a, g.T2, 8) + _masks(text.T, 2, 7, 7, 8)      # audio pass first, then the text pass

    sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() else v) for k, v in sd.items()}
    ma, mt = _pinned(masks, lambda: (O.encoder_forward(sdg, wave, tl)[0], O.encoder_forward_text(sdg, tokens, tok_len)[0]))
    ((ma * Ra).sum() + (mt * Rt).sum()).backward()
    ref = {k: v.grad for k, v in sdg.items() if v.is_floating_point() and v.grad is not None}
    ref["text_embed_tokens.weight"][1] = 0                          # padding_idx row (nn.Embedding(padding_idx) gives it no gradient)
    assert rel_l2(mem_t, mt.detach()) < 1e-5
    _compare(G, ref)
    assert "embed_bwd" in emu.calls


def test_fused_adam_host_logic_against_the_reference_goldens():
    """step_size = lr sqrt(1 - b2^t) / (1 - b1^t) and the device scalar block, formed by FusedAdam.advance (fairseq/optim/adam.py:197-224)."""
    g = np.load(os.path.join(GOLDEN, "adam.npz"))
    hp = dict(lr=float(g["lr"]), betas=tuple(float(x) for x in g["betas"]), eps=float(g["eps"]), weight_decay=float(g["weight_decay"]))
    params = {"w": torch.from_numpy(g["p0"]).clone(), "frozen": torch.ones(3)}
    opt = FusedAdam(params, lib=EmuLib(), **hp)
    for t in range(3):
        opt.advance()
        opt.step({"w": torch.from_numpy(g["grads"][t]).contiguous()})
        assert rel_l2(params["w"], torch.from_numpy(g["after"][t])) < 1e-6
    assert torch.equal(params["frozen"], torch.ones(3))             # no gradient: left alone
    # grad_scale (clipping coefficient / inverse loss scale) reaches the update
    ref = params["w"].clone()
    m, v = opt.m["w"].clone(), opt.v["w"].clone()
    gr = torch.from_numpy(g["grads"][0])
    opt.advance(lr=3e-4, grad_scale=0.25)
    opt.step({"w": gr.contiguous()})
    adam_oracle.adam_step(ref, gr * 0.25, m, v, 4, **dict(hp, lr=3e-4))
    assert rel_l2(params["w"], ref) < 1e-6


def test_split_k_weight_gradient_and_strided_window_operand():
    """_Ops.wgrad: with few output tiles and a long reduction the row axis becomes a GEMM batch (chunk layout of cst_transpose) whose
    partial products cst_colsum adds; `ldx < K` reads the overlapping windows of a strided convolution's input in place."""
    from chimera_st_b200.train import _Ops
    emu = EmuLib()
    o = _Ops(torch.device("cpu"), lib=emu)
    g = torch.Generator().manual_seed(4)
    rows, N, K = 3100, 64, 128
    dy, x = torch.randn(rows, N, generator=g), torch.randn(rows, K, generator=g)
    dW, dy_op = o.wgrad(dy, x, rows, N, K)
    assert dy_op is dy and rel_l2(dW, dy.T @ x) < 1e-6
    assert emu.calls.count("gemm") == 1 and "colsum" in emu.calls            # one batched launch (split-K 3) + the fixed-order reduction
    # conv window view: k = 2, stride 1 over 64-channel frames -> K = 128 columns at row pitch 64
    frames = torch.randn(rows + 1, 64, generator=g)
    dWc, _ = o.wgrad(dy, frames, rows, N, K, ldx=64)
    win = torch.cat((frames[:-1], frames[1:]), 1)
    assert rel_l2(dWc, dy.T @ win) < 1e-6
    # full linear backward with a residual gradient added to dx
    W, res = torch.randn(N, K, generator=g), torch.randn(rows, K, generator=g)
    dx, dW2, db = o.linear_bwd(x, W, dy, rows, dx_residual=res)
    assert rel_l2(dx, dy @ W + res) < 1e-6 and rel_l2(dW2, dy.T @ x) < 1e-6 and rel_l2(db, dy.sum(0)) < 1e-6


def test_emulated_dropout_replays_in_the_backward_pass():
    """Dropout of the training recipe (elementwise sites + attention probabilities): the forward's masks are regenerated (never stored)
    in the backward pass.  Autograd through the oracle with THE SAME masks installed at the reference's dropout sites gives the same
    memories and gradients."""
    torch.set_num_threads(8)
    lens = [4700, 3300]
    sd = synth.make_state_dict(seed=0, interlingua_length=8, dead_heads=False)
    wave, tl = synth.make_waveforms(lens, seed=31)
    R = torch.randn(8, 2, 512, generator=torch.Generator().manual_seed(1))
    emu = EmuLib()
    step = EncoderTrainStep(sd, 2, wave.shape[1], device="cpu", feature_grad_mult=1.0, lib=emu, dropout=0.1, activation_dropout=0.2,
                            attention_dropout=0.3, w2v_dropout=0.1, w2v_dropout_input=0.15, seed=77)
    g = step.g
    mem, G = step.forward_backward(wave, tl, R)
    n_elem, n_prob = 2 + 2 * 12 + 1 + 3 * 6 + 3 * 3, 12 + 6 + 3
    n_sites = n_elem + n_prob
    assert len(step._sites) == n_sites and emu.calls.count("dropout") == 2 * n_elem           # every site once forward, once backward
    assert emu.calls.count("attention_dropout_fwd") == n_prob == emu.calls.count("attention_bwd_tc_dropout")
    assert "attention" not in emu.calls and "attention_bwd" not in emu.calls
    fwd, bwd = emu.dropout_log[:n_elem], emu.dropout_log[n_elem:]
    assert sorted(fwd) == sorted(bwd)                                                         # same (seed, site, shape, p) both ways

    def p_of(tag):
        if tag.endswith(".prob"):
            return 0.1 if tag.startswith("w2v") else 0.3
        return 0.15 if tag == "w2v.input" else 0.2 if tag.endswith(".act") else 0.1

    def geom(tag):
        if tag.endswith(".prob"):
            return (g.T6a, g.Tp) if tag.startswith("w2v") else (8, g.T2) if tag.startswith("mem") else (g.T2a, g.T2)
        return g.T6a if tag.startswith("w2v") else 8 if tag.startswith("mem") else g.T2a
    hook, used = oracle_dropout_hook(step, 0, geom, p_of)
    O.DROPOUT_HOOK = hook
    try:
        ref_mem, ref = _autograd(sd, lambda s: O.encoder_forward(s, wave, tl)[0], R, _masks(step.T, g.B, g.T2a, g.T2, 8))
    finally:
        O.DROPOUT_HOOK = None
    assert len(set(used)) == n_sites
    assert rel_l2(mem, ref_mem) < 1e-5
    _compare(G, ref)
    # it IS dropout: the memories differ from the dropout-free ones, and a new seed draws new masks
    mem0 = EncoderTrainStep(sd, 2, wave.shape[1], device="cpu", lib=EmuLib()).forward(wave, tl)
    assert rel_l2(mem, mem0) > 1e-2
    step.next_dropout_seed()
    assert int(step.seed_dev.item()) == 78
    assert rel_l2(step.forward(wave, tl), mem) > 1e-2


def test_emulated_text_pass_draws_its_own_dropout_masks():
    """The text pass of a step shares the layers with the audio pass but not the masks: its sites are numbered separately (pass 1), and
    replaying them in the oracle's text branch reproduces memories and gradients."""
    torch.set_num_threads(8)
    V = 60
    sd = synth.make_state_dict(seed=2, interlingua_length=8, dead_heads=False, text_vocab=V)
    gen = torch.Generator().manual_seed(9)
    tok_len = torch.tensor([7, 4])
    tokens = torch.randint(4, V, (2, 7), generator=gen)
    tokens[1, 4:] = 1
    Rt = torch.randn(8, 2, 512, generator=gen)
    wave, tl = synth.make_waveforms([4100], seed=5)
    emu = EmuLib()
    step = EncoderTrainStep(sd, 1, wave.shape[1], device="cpu", feature_grad_mult=1.0, lib=emu, dropout=0.2, seed=3)
    text = TextTrainPass(step, 2, 7)
    step.forward(wave, tl)                                          # audio pass first: registers the pass-0 sites
    n_audio = len(step._sites)
    mem_t = text.forward(tokens, tok_len)
    G = text.backward(Rt)
    n_text = 1 + 4 * 6 + 4 * 3                                      # embed + (attn, prob, act, ffn) x (6 shared + 3 memory layers)
    assert len(step._sites) == n_audio + n_text
    assert all(step._sites[(1, t)] != step._sites[(0, t)] for t in ("embed", "enc0.attn", "mem2.prob"))

    def geom(tag):
        if tag.endswith(".prob"):
            return (8, 7) if tag.startswith("mem") else (7, 7)
        return 8 if tag.startswith("mem") else 7
    hook, used = oracle_dropout_hook(step, 1, geom, lambda tag: 0.2)
    sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() else v) for k, v in sd.items()}
    O.DROPOUT_HOOK = hook
    try:
        mt = _pinned(_masks(text.T, 2, 7, 7, 8), lambda: O.encoder_forward_text(sdg, tokens, tok_len)[0])
    finally:
        O.DROPOUT_HOOK = None
    (mt * Rt).sum().backward()
    ref = {k: v.grad for k, v in sdg.items() if v.is_floating_point() and v.grad is not None}
    ref["text_embed_tokens.weight"][1] = 0
    assert len(set(used)) == n_text
    assert rel_l2(mem_t, mt.detach()) < 1e-5
    _compare(G, ref)


def test_emulated_training_update_end_to_end():
    """One whole update on CPU: gradients (emulated launch sequence) -> FusedAdam at the schedule's learning rate -> refresh of the
    kernel-layout operands -> next forward.  Reference: autograd through the oracle -> the oracle restatement of fairseq's Adam (pinned
    bit-exactly to the reference class, tests/test_adam.py) -> oracle forward with the updated state dict."""
    from chimera_st_b200.train import InverseSqrtLR
    torch.set_num_threads(8)
    sd = synth.make_state_dict(seed=0, interlingua_length=8, dead_heads=False)
    wave, tl = synth.make_waveforms([4300, 3100], seed=12)
    R = torch.randn(8, 2, 512, generator=torch.Generator().manual_seed(2))
    emu = EmuLib()
    # the step keeps the master parameters it is given (no copy on the same device): hand it its own tensors
    step = EncoderTrainStep({k: v.clone() for k, v in sd.items()}, 2, wave.shape[1], device="cpu", feature_grad_mult=0.1, lib=emu)
    g = step.g
    step.forward(wave, tl)
    G = step.backward(R)
    masks = _masks(step.T, g.B, g.T2a, g.T2, 8)
    lr = InverseSqrtLR(1e-2, 10).at(5)                              # mid warm-up: 5e-3
    hp = dict(betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-4)
    opt = FusedAdam({k: step.sd[k] for k in G}, lr=1.0, lib=emu, **hp)
    opt.advance(lr=lr)
    opt.step({k: v.contiguous() for k, v in G.items()})
    step.refresh_weights()
    mem1 = step.forward(wave, tl)
    # reference update
    _, ref = _autograd(sd, lambda s: O.encoder_forward(s, wave, tl)[0], R, masks)
    sd_ref = {k: v.clone() for k, v in sd.items()}
    flips = total = 0
    for k, gr in ref.items():
        if float(gr.abs().max()) == 0 and k not in G:
            continue
        gr = gr * (0.1 if ".feature_extractor." in k else 1.0)     # GradMultiply(0.1) (wav2vec2.py:530-532)
        p = sd_ref[k]
        adam_oracle.adam_step(p, gr, torch.zeros_like(p), torch.zeros_like(p), 1, lr=lr, **hp)
        # the first Adam step moves every element by ~lr * sign(g): elements whose gradient is rounding noise may go either way
        d = (step.sd[k].reshape(p.shape) - p).abs() > 0.5 * lr
        flips += int(d.sum()); total += p.numel()
    assert flips < 5e-4 * total, (flips, total)                     # measured 5.7e-5
    mem1_ref, _ = O.encoder_forward(sd_ref, wave, tl)
    # the updated operands reached the kernels: the new memories follow the reference's (far from the old ones, close to the new)
    mem0_ref, _ = O.encoder_forward(sd, wave, tl)
    assert rel_l2(mem1, mem0_ref) > 20 * rel_l2(mem1, mem1_ref)
    assert rel_l2(mem1, mem1_ref) < 1e-3                            # measured 5.9e-5 (against 5.06 to the old memories)
