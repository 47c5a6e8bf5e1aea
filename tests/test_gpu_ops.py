"""GPU parity of the individual C-ABI kernels against CPU restatements (torch fp32/fp64 on host)."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import chimera_st_b200  # noqa: F401
from chimera_st_b200 import synth, lengths
from oracle import chimera_oracle as O
from conftest import rel_l2, rel_max, GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda"


def ops():
    from chimera_st_b200 import ops as _ops
    return _ops


def L_():
    from chimera_st_b200 import _lib
    return _lib


def test_device_is_blackwell():
    name, sms, major, minor = L_().device_info()
    assert major == 10, (name, major, minor)


def test_frame_lengths_exhaustive_vs_reference_rows():
    rows = np.load(os.path.join(GOLDEN, "lengths.npz"))["rows"]
    by_L = {}
    for Lw, n, T, v, s2 in rows.tolist():
        by_L.setdefault(Lw, []).append((n, T, v, s2))
    for Lw, items in by_L.items():
        lens = torch.tensor([n for n, *_ in items], dtype=torch.int64, device=DEV)
        w2v, sub, l64, mask = ops().frame_lengths(lens, Lw)
        assert w2v.tolist() == [v for _, _, v, _ in items], Lw
        assert sub.tolist() == [s for *_, s in items], Lw
        assert l64.tolist() == w2v.tolist()
        ref_mask = O.frame_padding_mask(lens.cpu(), items[0][1])
        assert torch.equal(mask.cpu(), ref_mask), Lw


@pytest.mark.parametrize("lens", [[16000, 12345, 8000], [4000], [80000, 64000, 48123, 32000]])
def test_conv0_groupnorm_gelu(lens):
    sd = synth.make_state_dict(seed=0)
    wave, _ = synth.make_waveforms(lens, seed=3)
    wave[0] += 0.05                                     # DC offset: stresses the variance-from-moments path
    P = "wav2vec_model.feature_extractor.conv_layers."
    ref = O.conv_feature_extractor(sd, wave, upto=1).transpose(1, 2)          # [B,T0,512]
    T0 = ref.shape[1]
    rps = 64 * ((T0 + 63) // 64)
    out, _ = ops().conv0_gn_gelu(wave.to(DEV), sd[P + "0.0.weight"].to(DEV), sd[P + "0.2.weight"].to(DEV),
                                 sd[P + "0.2.bias"].to(DEV), torch.float32, rps)
    out = out.cpu()
    assert rel_l2(out[:, :T0], ref) < 2e-6 and rel_max(out[:, :T0], ref) < 2e-5
    assert float(out[:, T0:].abs().max()) == 0.0 if rps > T0 else True
    outb, _ = ops().conv0_gn_gelu(wave.to(DEV), sd[P + "0.0.weight"].to(DEV), sd[P + "0.2.weight"].to(DEV),
                                  sd[P + "0.2.bias"].to(DEV), torch.bfloat16, rps)
    assert rel_l2(outb.cpu().float()[:, :T0], ref) < 4e-3      # bf16 rounding of the stored value only
    # tensor-core conv0 (3-term fp16 split): must agree with the CUDA-core kernel to the output rounding, zero the tail
    for dt, tol in ((torch.float16, 5e-4), (torch.bfloat16, 4e-3)):
        outc, _ = ops().conv0_gn_gelu(wave.to(DEV), sd[P + "0.0.weight"].to(DEV), sd[P + "0.2.weight"].to(DEV),
                                      sd[P + "0.2.bias"].to(DEV), dt, rps)
        outt, _ = ops().conv0_gn_gelu(wave.to(DEV), sd[P + "0.0.weight"].to(DEV), sd[P + "0.2.weight"].to(DEV),
                                      sd[P + "0.2.bias"].to(DEV), dt, rps, tensor_core=True)
        outt = outt.cpu().float()
        assert rel_l2(outt[:, :T0], ref) < tol
        assert float(outt[:, T0:].abs().max()) == 0.0 if rps > T0 else True
        # same value before rounding up to ~1e-6: at most a 1-ulp flip on a tiny fraction of the outputs
        diff = (outt - outc.cpu().float()).abs()
        assert float((diff > 0).float().mean()) < 0.02
        assert rel_l2(outt[:, :T0], outc.cpu().float()[:, :T0]) < (2e-4 if dt == torch.float16 else 1.5e-3)


def _ref_gemm(A, W, bias, act, alpha, residual):
    acc = A.double() @ W.double().T
    if bias is not None:
        acc = acc + bias.double()
    if act == 1:
        acc = 0.5 * acc * (1 + torch.erf(acc / math.sqrt(2.0)))
    elif act == 2:
        acc = torch.relu(acc)
    elif act == 3:
        acc = acc[:, 0::2] * torch.sigmoid(acc[:, 1::2])
    acc = acc * alpha
    if residual is not None:
        acc = acc + residual.double()
    return acc


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 2e-5)])
@pytest.mark.parametrize("M,N,K,act", [(300, 768, 512, 0), (257, 512, 1536, 1), (1000, 2304, 768, 0),
                                       (128, 3072, 768, 1), (513, 768, 3072, 0), (200, 1024, 3840, 3),
                                       (64, 2048, 512, 2), (129, 512, 2048, 0), (100, 1536, 512, 0)])
def test_linear(dtype, tol, M, N, K, act):
    """Both GEMM kernels (FFMA fp32 / tcgen05 bf16) on identical operand values: the bf16 run gets
    bf16-rounded inputs, so the only difference from the fp64 reference is fp32 accumulation order."""
    g = torch.Generator().manual_seed(M * 7 + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(dtype)
    W = (torch.randn(N, K, generator=g) * (1.0 / math.sqrt(K))).to(dtype)
    bias = torch.randn(N, generator=g) * 0.1
    n_out = N // 2 if act == 3 else N
    res = torch.randn(M, n_out, generator=g)
    ref = _ref_gemm(A.float(), W.float(), bias, act, 1.5, res)
    out = ops().linear(A.to(DEV), W.to(DEV), bias.to(DEV), act=act, residual=res.to(DEV), alpha=1.5,
                       out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert rel_l2(out.cpu(), ref) < tol, rel_l2(out.cpu(), ref)
    assert rel_max(out.cpu(), ref) < 20 * tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 2e-5), (torch.float16, 2e-5)])
def test_implicit_conv_gemm_matches_conv1d(dtype, tol):
    """stride-2 conv over channels-last rows as an overlapping-window GEMM (lda < K)."""
    g = torch.Generator().manual_seed(5)
    B, Tin, Cc, k = 3, 400, 512, 3
    x = (torch.randn(B, Tin, Cc, generator=g) * 0.5).to(dtype)
    w = (torch.randn(512, Cc, k, generator=g) * 0.03).to(dtype)
    ref = F.gelu(F.conv1d(x.float().transpose(1, 2).double(), w.float().double(), stride=2)).transpose(1, 2)  # [B,Tout,512]
    Tout = ref.shape[1]
    Ta = Tin // 2
    xb = torch.zeros(B * Tin + 8, Cc, dtype=dtype)
    xb[:B * Tin] = x.reshape(B * Tin, Cc)
    wk = w.permute(0, 2, 1).reshape(512, k * Cc).contiguous()
    out = torch.zeros(B * Ta, 512, dtype=torch.float32, device=DEV)
    ops().gemm(xb.to(DEV), wk.to(DEV), out, B * Ta, 512, k * Cc, lda=2 * Cc, a_rows=(B * Tin + 8) // 2, act=1,
               rows_per_seg=Ta)
    got = out.cpu().view(B, Ta, 512)[:, :Tout]
    assert rel_l2(got, ref) < tol, rel_l2(got, ref)
    if dtype != torch.float32:          # 16-bit output of the same dtype (what the conv stack stores between layers)
        out16 = torch.zeros(B * Ta, 512, dtype=dtype, device=DEV)
        ops().gemm(xb.to(DEV), wk.to(DEV), out16, B * Ta, 512, k * Cc, lda=2 * Cc, a_rows=(B * Tin + 8) // 2, act=1,
                   rows_per_seg=Ta)
        got16 = out16.cpu().float().view(B, Ta, 512)[:, :Tout]
        assert rel_l2(got16, ref) < (4e-3 if dtype == torch.bfloat16 else 5e-4)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 2e-5)])
def test_gemm_row_remap_mask_and_batches(dtype, tol):
    """segment remap (padded destination), per-segment zeroing and the (outer, inner) batch of the pos-conv."""
    g = torch.Generator().manual_seed(9)
    B, G, T, Kc = 2, 16, 70, 128 * 64
    Tpp = T + 128
    xg = (torch.randn(B * G * Tpp + 8, 64, generator=g) * 0.3).to(dtype)
    w = (torch.randn(G, 48, Kc, generator=g) * 0.01).to(dtype)
    bias = torch.randn(768, generator=g) * 0.1
    res = torch.randn(B * T, 768, generator=g)
    out = torch.zeros(B * T, 768, dtype=torch.float32, device=DEV)
    ops().gemm(xg.to(DEV), w.to(DEV), out, T, 48, Kc, lda=64, a_rows=Tpp, bias=bias.to(DEV), residual=res.to(DEV),
               act=1, ldc=768, nb_outer=B, nb_inner=G, a_bs=(G * Tpp * 64, Tpp * 64), w_bs=48 * Kc,
               c_bs=(T * 768, 48), bias_bs=48)
    ref = torch.zeros(B, T, 768, dtype=torch.float64)
    flat = xg.float().double().reshape(-1)
    for b in range(B):
        for gi in range(G):
            base = (b * G + gi) * Tpp * 64
            Am = flat[base:base + Tpp * 64].as_strided((T, Kc), (64, 1))
            acc = Am @ w[gi].float().double().T + bias[gi * 48:(gi + 1) * 48].double()
            ref[b, :, gi * 48:(gi + 1) * 48] = 0.5 * acc * (1 + torch.erf(acc / math.sqrt(2.0)))
    ref = ref.view(B * T, 768) + res.double()
    assert rel_l2(out.cpu(), ref) < tol, rel_l2(out.cpu(), ref)
    # remap + mask: 3 segments of 50 rows, 45 valid, written at offset 2 into 60-row segments, zero beyond seg_len
    M, N, K = 150, 512, 512
    A = (torch.randn(M, K, generator=g) * 0.5).to(dtype)
    W = (torch.randn(N, K, generator=g) * 0.05).to(dtype)
    seg_len = torch.tensor([45, 20, 33], dtype=torch.int32)
    out = torch.full((3 * 60, N), 7.0, dtype=torch.float32, device=DEV)
    ops().gemm(A.to(DEV), W.to(DEV), out, M, N, K, lda=K, a_rows=M, rows_per_seg=50, seg_rows_valid=45,
               out_rows_per_seg=60, out_row_off=2, seg_len=seg_len.to(DEV))
    full = (A.float().double() @ W.float().double().T).view(3, 50, N)
    exp = torch.full((3, 60, N), 7.0, dtype=torch.float64)
    for s in range(3):
        exp[s, 2:2 + 45] = full[s, :45]
        exp[s, 2 + int(seg_len[s]):2 + 45] = 0
    assert rel_l2(out.cpu().view(3, 60, N), exp) < tol


@pytest.mark.parametrize("Cd", [512, 768])
def test_layernorm(Cd):
    g = torch.Generator().manual_seed(Cd)
    x = torch.randn(301, Cd, generator=g) * 3 + 1
    gm, bt = 1 + 0.1 * torch.randn(Cd, generator=g), 0.1 * torch.randn(Cd, generator=g)
    ref = F.layer_norm(x.double(), (Cd,), gm.double(), bt.double(), 1e-5)
    o32, obf = ops().layernorm(x.to(DEV), gm.to(DEV), bt.to(DEV), torch.bfloat16)
    assert rel_l2(o32.cpu(), ref) < 1e-6
    assert rel_l2(obf.cpu().float(), ref) < 4e-3


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 6e-3)])
@pytest.mark.parametrize("B,H,Tq,Tk,masked", [(3, 12, 249, 249, True), (2, 8, 63, 63, True), (2, 8, 16, 63, False),
                                              (1, 12, 130, 130, True), (2, 8, 64, 300, False),
                                              (2, 12, 1499, 1499, True), (3, 8, 188, 188, True), (2, 12, 750, 749, False),
                                              (1, 12, 128, 128, False), (2, 8, 129, 257, True), (3, 12, 400, 385, True),
                                              (3, 8, 16, 375, True), (3, 8, 40, 701, True), (2, 8, 64, 1500, False),
                                              (2, 8, 16, 3100, False)])
def test_attention(dtype, tol, B, H, Tq, Tk, masked):
    g = torch.Generator().manual_seed(Tq * 3 + Tk)
    Cd = H * 64
    q = (torch.randn(B, Tq, Cd, generator=g) * 0.5).to(dtype)
    k = torch.randn(B, Tk, Cd, generator=g).to(dtype)
    v = torch.randn(B, Tk, Cd, generator=g).to(dtype)
    kl = torch.tensor([Tk, max(1, Tk // 2), 1][:B], dtype=torch.int32) if masked else None
    qh = q.float().double().view(B, Tq, H, 64).transpose(1, 2)
    kh = k.float().double().view(B, Tk, H, 64).transpose(1, 2)
    vh = v.float().double().view(B, Tk, H, 64).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2)
    if masked:
        s = s.masked_fill(torch.arange(Tk)[None, None, None, :] >= kl.long()[:, None, None, None], float("-inf"))
    ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, Tq, Cd)
    out = ops().attention(q.to(DEV), k.to(DEV), v.to(DEV), H, kl.to(DEV) if masked else None)
    assert rel_l2(out.cpu().float(), ref) < tol, rel_l2(out.cpu().float(), ref)


@pytest.mark.parametrize("M,N,K,lazy_res,inplace", [(300, 768, 768, True, True), (300, 768, 3072, True, True), (20000, 768, 768, True, True),
                                                    (517, 512, 512, False, True), (130, 768, 768, False, False), (16500, 512, 2048, False, True)])
def test_gemm_fused_layernorm_producer(M, N, K, lazy_res, inplace):
    """out-proj / fc2 with the LayerNorm fused around the GEMM: y <- LN(y_prev)*g + b (from partial statistics, or the plain
    residual) + A W^T + bias, in place; a bf16 copy; partial statistics of the new rows.  Checked against torch in fp64 on
    the bf16-rounded operands (single-CTA and CTA-pair kernels: M on both sides of 16384)."""
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.7).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * 0.03).to(torch.bfloat16)
    bias = torch.randn(N, generator=g) * 0.1
    yprev = torch.randn(M, N, generator=g) * 1.7 + 0.4
    gamma, beta = 1 + 0.2 * torch.randn(N, generator=g), 0.1 * torch.randn(N, generator=g)
    slots = 2 * (N // 256)
    # partial statistics of yprev as a producer would have written them: per 128-column slice {sum, sum of squares}
    st_in = torch.zeros(M, 8, 2)
    for sidx in range(slots):
        blk = yprev[:, 128 * sidx:128 * (sidx + 1)].double()
        st_in[:, sidx, 0], st_in[:, sidx, 1] = blk.sum(1).float(), (blk * blk).sum(1).float()
    res = torch.nn.functional.layer_norm(yprev.double(), (N,), gamma.double(), beta.double(), 1e-5) if lazy_res else yprev.double()
    ref = A.float().double() @ W.float().double().T + bias.double() + res
    yd = yprev.to(DEV).clone()
    out = yd if inplace else torch.empty(M, N, device=DEV)
    c2 = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    st_out = torch.full((M, 8, 2), -1.0, device=DEV)
    ops().gemm(A.to(DEV), W.to(DEV), out, M, N, K, lda=K, a_rows=M, bias=bias.to(DEV), residual=yd,
               res_ln=(st_in.to(DEV), slots, gamma.to(DEV), beta.to(DEV)) if lazy_res else None, c2=c2, out_stats=st_out, ln_dim=N)
    torch.cuda.synchronize()
    got = out.cpu().double()
    assert rel_l2(got, ref) < 5e-6, rel_l2(got, ref)
    assert torch.equal(c2.cpu(), out.cpu().to(torch.bfloat16))
    so = st_out.cpu().double()
    assert float((so[:, :slots, 0].sum(1) - got.sum(1)).abs().max()) < 2e-3
    assert float(((so[:, :slots, 1].sum(1) - (got * got).sum(1)) / (got * got).sum(1)).abs().max()) < 2e-6
    for sidx in range(slots):                               # every slot is one 128-column slice
        assert float((so[:, sidx, 0] - got[:, 128 * sidx:128 * (sidx + 1)].sum(1)).abs().max()) < 1e-3
    assert bool((so[:, slots:] == -1.0).all())


@pytest.mark.parametrize("M,N,K,act", [(300, 2304, 768, 0), (20000, 3072, 768, 1), (517, 1536, 512, 0), (130, 2048, 512, 2)])
def test_gemm_fused_layernorm_consumer(M, N, K, act):
    """QKV / fc1 on the bf16 copy of UN-normalised rows, LayerNorm applied after the product from the partial statistics:
    rstd*(y W'^T) - rstd*mean*colsum(W') + c  ==  act(LN(y; g, b) W^T + bias)."""
    g = torch.Generator().manual_seed(M + N)
    y = torch.randn(M, K, generator=g) * 1.5 + 0.3
    W = torch.randn(N, K, generator=g) * 0.04
    bias = torch.randn(N, generator=g) * 0.1
    gamma, beta = 1 + 0.2 * torch.randn(K, generator=g), 0.1 * torch.randn(K, generator=g)
    slots = 2 * (K // 256)
    st = torch.zeros(M, 8, 2)
    for sidx in range(slots):
        blk = y[:, 128 * sidx:128 * (sidx + 1)].double()
        st[:, sidx, 0], st[:, sidx, 1] = blk.sum(1).float(), (blk * blk).sum(1).float()
    Wf = (W.double() * gamma.double()[None, :]).float().to(torch.bfloat16)
    cs = Wf.double().sum(1).float()
    cb = (bias.double() + W.double() @ beta.double()).float()
    ya = y.to(torch.bfloat16)
    # what the kernel computes, in fp64, on the rounded operands
    mean, var = y.double().mean(1, keepdim=True), y.double().var(1, unbiased=False, keepdim=True)
    rstd = (var + 1e-5).rsqrt()
    z = rstd * (ya.float().double() @ Wf.float().double().T) - rstd * mean * cs.double()[None, :] + cb.double()[None, :]
    ref = {0: z, 1: torch.nn.functional.gelu(z), 2: torch.relu(z)}[act]
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops().gemm(ya.to(DEV), Wf.to(DEV), out, M, N, K, lda=K, a_rows=M, bias=cb.to(DEV), act=act,
               ln_in=(st.to(DEV), cs.to(DEV), slots), ln_dim=K)
    torch.cuda.synchronize()
    assert rel_l2(out.cpu().float(), ref) < 3e-3, rel_l2(out.cpu().float(), ref)
    # and it IS the LayerNorm'ed projection (up to bf16 rounding of operands / output)
    true = torch.nn.functional.layer_norm(y.double(), (K,), gamma.double(), beta.double(), 1e-5) @ W.double().T + bias.double()
    true = {0: true, 1: torch.nn.functional.gelu(true), 2: torch.relu(true)}[act]
    assert rel_l2(out.cpu().float(), true) < 8e-3, rel_l2(out.cpu().float(), true)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 6e-3)])
@pytest.mark.parametrize("H,segs,masked", [
    (12, [(749, 749), (300, 290), (1499, 1499), (128, 100), (257, 256)], True),      # wav2vec2 stage: q rows >= kv rows
    (8, [(188, 188), (63, 63), (375, 375), (1, 1), (26, 25)], True),                 # shared layers
    (8, [(16, 188), (16, 63), (16, 375), (16, 25)], False),                          # memory stage: M queries, no mask
    (8, [(64, 250), (64, 1000)], False)])
def test_attention_segment_table(dtype, tol, H, segs, masked):
    """cst_attention_segs: utterances of different lengths in one row space (the super-batch form), q / kv rows of
    utterance b anywhere in their buffers; must equal per-utterance softmax attention in fp64."""
    g = torch.Generator().manual_seed(len(segs) * 7 + H)
    Cd = H * 64
    gap = 3                                                         # junk rows between segments: must never be read as valid
    tbl, qo, ko = [], 0, 0
    for nq, nk in segs:
        tbl.append([qo, nq, ko, nk])
        qo += nq + gap
        ko += nk + gap
    q = (torch.randn(qo, Cd, generator=g) * 0.5).to(dtype)
    k = torch.randn(ko, Cd, generator=g).to(dtype)
    v = torch.randn(ko, Cd, generator=g).to(dtype)
    kl = torch.tensor([max(1, (nk * (3 + i)) // (4 + i)) for i, (_, nk) in enumerate(segs)], dtype=torch.int32) if masked else None
    out = torch.full((qo, Cd), 7.0, dtype=dtype, device=DEV)
    L = L_()
    seg = torch.tensor(tbl, dtype=torch.int32, device=DEV)
    qd, kd, vd = q.to(DEV), k.to(DEV), v.to(DEV)
    kld = kl.to(DEV) if masked else None
    L.check(L.load().cst_attention_segs(qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), out.data_ptr(), L.DT[dtype], Cd, Cd, Cd,
                                        len(segs), H, seg.data_ptr(), max(s[0] for s in segs), max(s[1] for s in segs), qo, ko,
                                        L.ptr(kld), L.stream_ptr()))
    torch.cuda.synchronize()
    out = out.cpu().float()
    for i, (q0, nq, k0, nk) in enumerate(tbl):
        qh = q[q0:q0 + nq].float().double().view(nq, H, 64).transpose(0, 1)
        kh = k[k0:k0 + nk].float().double().view(nk, H, 64).transpose(0, 1)
        vh = v[k0:k0 + nk].float().double().view(nk, H, 64).transpose(0, 1)
        sc = qh @ kh.transpose(-1, -2)
        if masked:
            sc = sc.masked_fill(torch.arange(nk)[None, None, :] >= int(kl[i]), float("-inf"))
        ref = (torch.softmax(sc, -1) @ vh).transpose(0, 1).reshape(nq, Cd)
        assert rel_l2(out[q0:q0 + nq], ref) < tol, (i, rel_l2(out[q0:q0 + nq], ref))
        if i + 1 < len(tbl):
            assert bool((out[q0 + nq:q0 + nq + gap] == 7.0).all()), "rows outside the segment were written"


@pytest.mark.parametrize("B,T", [(2, 70), (3, 250), (1, 129), (2, 750), (1, 1000), (1, 385)])
def test_resident_posconv_matches_generic_gemm_path(B, T):
    """cst_posconv (panel resident in smem, row-shifted swizzled A descriptors) vs the batched implicit GEMM."""
    g = torch.Generator().manual_seed(B * 100 + T)
    G, Kc = 16, 128 * 64
    Tpp = T + 128
    x = torch.randn(B * T, 768, generator=g)
    w = (torch.randn(G, 48, Kc, generator=g) * 0.02).to(torch.bfloat16)
    bias = torch.randn(768, generator=g) * 0.1
    xd = x.to(DEV)
    xg = torch.zeros(B * G * Tpp + 8, 64, dtype=torch.bfloat16, device=DEV)
    L = L_()
    L.check(L.load().cst_posconv_pack(xd.data_ptr(), B, T, T, xg.data_ptr(), L.BF16, Tpp, L.stream_ptr()))
    ref = torch.zeros(B * T, 768, dtype=torch.float32, device=DEV)
    ops().gemm(xg, w.to(DEV), ref, T, 48, Kc, lda=64, a_rows=Tpp, bias=bias.to(DEV), residual=xd, act=1, ldc=768,
               nb_outer=B, nb_inner=G, a_bs=(G * Tpp * 64, Tpp * 64), w_bs=48 * Kc, c_bs=(T * 768, 48), bias_bs=48)
    out = torch.zeros(B * T, 768, dtype=torch.float32, device=DEV)
    wd, bd = w.to(DEV), bias.to(DEV)
    L.check(L.load().cst_posconv(xg.data_ptr(), wd.data_ptr(), bd.data_ptr(), xd.data_ptr(), out.data_ptr(), B, T, T, Tpp,
                                 L.stream_ptr()))
    torch.cuda.synchronize()
    assert rel_l2(out.cpu(), ref.cpu()) < 2e-5, rel_l2(out.cpu(), ref.cpu())
    # "taps stacked in M" formulation: weights as the A operand, two taps per instruction (cst_posconv_stacked)
    w4 = w.view(G, 48, 128, 64)                                        # [g, co, tap, lane]
    w2 = torch.zeros(G, 64, 128, 64, dtype=torch.bfloat16)
    w2[:, :, 0:48, :] = w4[:, :, 0::2, :].permute(0, 2, 1, 3)
    w2[:, :, 64:112, :] = w4[:, :, 1::2, :].permute(0, 2, 1, 3)
    out2 = torch.zeros(B * T, 768, dtype=torch.float32, device=DEV)
    w2d = w2.to(DEV).contiguous()
    L.check(L.load().cst_posconv_stacked(xg.data_ptr(), w2d.data_ptr(), bd.data_ptr(), xd.data_ptr(), out2.data_ptr(), B, T, T, Tpp,
                                         L.stream_ptr()))
    torch.cuda.synchronize()
    assert rel_l2(out2.cpu(), ref.cpu()) < 2e-5, rel_l2(out2.cpu(), ref.cpu())


def test_conv0_tensor_core_full_c2_size_matches_cuda_core_kernel():
    """B = 32 x 15 s (BASELINE configs[1] size): 12 000 frame tiles over the persistent grid, utterance changes inside a CTA's
    tile range, dead frames in each utterance's last tile -- against the CUDA-core kernel on the same inputs."""
    sd = synth.make_state_dict(seed=0)
    lens = [240000 - 997 * i for i in range(32)]
    wave, _ = synth.make_waveforms(lens, seed=11)
    P = "wav2vec_model.feature_extractor.conv_layers."
    T0 = (wave.shape[1] - 10) // 5 + 1
    rps = 64 * ((T0 + 63) // 64)
    args = (wave.to(DEV), sd[P + "0.0.weight"].to(DEV), sd[P + "0.2.weight"].to(DEV), sd[P + "0.2.bias"].to(DEV), torch.float16, rps)
    a, _ = ops().conv0_gn_gelu(*args)
    b, _ = ops().conv0_gn_gelu(*args, tensor_core=True)
    assert float(b[:, T0:].abs().max()) == 0.0 if rps > T0 else True
    diff = (a.float() - b.float()).abs()
    assert float((diff > 0).float().mean()) < 0.02                      # 1-ulp flips only
    assert float(diff.max()) <= 2.0 ** -9 * float(a.float().abs().max())
    assert rel_l2(b.float().cpu(), a.float().cpu()) < 2e-4
