"""TEST INFRASTRUCTURE: a host (numpy/torch) emulator of the C ABI in include/chimera_st_b200.h.

It interprets the same argument structs / pointers as the CUDA library, on HOST memory, in fp32.
Purpose: run the real `EncoderPlan` launch sequence on CPU tensors so that all host-side logic --
frame geometry, buffer layouts, weight re-arrangement, row remapping, the order of launches -- is
checked against the oracle without a GPU.  It doubles as an executable statement of each entry
point's semantics.  Never imported by the product package.
"""
import ctypes as C
import math

import numpy as np
import torch

F32 = 0


def _mem(ptr, n, dtype=np.float32):
    if n <= 0:
        return np.zeros(0, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype)


class EmuLib:
    def __init__(self):
        self.calls = []
        self.dropout_log = []

    def cst_wave_i16_to_f32(self, src, dst, n, stream):
        _mem(dst, n)[:] = _mem(src, n, np.int16).astype(np.float32) * np.float32(1.0 / 32768.0)
        self.calls.append("wave_i16_to_f32")
        return 0

    # ---- a4
    def cst_frame_lengths(self, src_len, B, L, n_frames, w2v_valid, sub_valid, w2v_len64, frame_mask, stream):
        lens = _mem(src_len, B, np.int64)
        r = L // n_frames
        v = np.minimum(n_frames, -(-lens // r))
        if w2v_valid:
            _mem(w2v_valid, B, np.int32)[:] = v
        if w2v_len64:
            _mem(w2v_len64, B, np.int64)[:] = v
        if sub_valid:
            _mem(sub_valid, B, np.int32)[:] = ((v + 1) // 2 + 1) // 2
        if frame_mask:
            _mem(frame_mask, B * n_frames, np.uint8).reshape(B, n_frames)[:] = np.arange(n_frames)[None, :] >= v[:, None]
        self.calls.append("frame_lengths")
        return 0

    # ---- a1
    def cst_conv0_stats(self, wave, B, L, w, gamma, beta, scale_shift, ws, stream):
        x = torch.from_numpy(_mem(wave, B * L).reshape(B, L).copy()).double()
        wt = torch.from_numpy(_mem(w, 5120).reshape(512, 10).copy()).double()
        T0 = (L - 10) // 5 + 1
        win = x.unfold(1, 10, 5)[:, :T0]                              # [B,T0,10]
        S = win.sum(1)                                               # lag sums
        R = torch.einsum("bti,btj->bij", win, win)                   # lag correlations
        mean = (S @ wt.T) / T0
        ey2 = torch.einsum("ci,bij,cj->bc", wt, R, wt) / T0
        rstd = 1.0 / torch.sqrt((ey2 - mean * mean).clamp_min(0) + 1e-5)
        g = torch.from_numpy(_mem(gamma, 512).copy()).double()
        bt = torch.from_numpy(_mem(beta, 512).copy()).double()
        sc = g * rstd
        out = torch.stack((sc, bt - mean * sc), -1).float().numpy()
        _mem(scale_shift, B * 512 * 2).reshape(B, 512, 2)[:] = out
        self.calls.append("conv0_stats")
        return 0

    def cst_conv0_apply(self, wave, B, L, w, scale_shift, out, out_dtype, rows_per_seg, stream):
        assert out_dtype == F32
        x = torch.from_numpy(_mem(wave, B * L).reshape(B, L).copy())
        wt = torch.from_numpy(_mem(w, 5120).reshape(512, 1, 10).copy())
        ss = torch.from_numpy(_mem(scale_shift, B * 1024).reshape(B, 512, 2).copy())
        y = torch.nn.functional.conv1d(x.unsqueeze(1), wt, stride=5)  # [B,512,T0]
        y = y * ss[:, :, 0:1] + ss[:, :, 1:2]
        y = 0.5 * y * (1 + torch.erf(y / math.sqrt(2.0)))
        T0 = y.shape[2]
        o = _mem(out, B * rows_per_seg * 512).reshape(B, rows_per_seg, 512)
        o[:, :T0] = y.transpose(1, 2).numpy()
        o[:, T0:] = 0
        self.calls.append("conv0_apply")
        return 0

    # ---- GEMM
    def cst_gemm(self, pref, stream):
        p = pref._obj
        assert p.ab_dtype == F32 and p.c_dtype == F32
        assert p.N % 8 == 0 and p.K % 64 == 0 and p.lda % 8 == 0 and p.ldc % 8 == 0
        a_lim = p.a_rows * p.lda
        glu = p.act == 3
        n_out = p.N // 2 if glu else p.N
        max_row = ((p.M - 1) // p.rows_per_seg) * p.out_rows_per_seg + p.rows_per_seg + p.out_row_off
        for zo in range(p.nb_outer):
            for zi in range(p.nb_inner):
                a_off = zo * p.a_bs_outer + zi * p.a_bs_inner
                A = np.concatenate((_mem(p.A + 4 * a_off, a_lim), np.zeros(p.K + p.lda, np.float32)))
                At = torch.from_numpy(A.copy())
                need = (p.M - 1) * p.lda + p.K
                assert need <= At.numel()
                # rows whose window leaves the addressable region read zeros from there on (guarded loads)
                Am = At.as_strided((p.M, p.K), (p.lda, 1))
                W = torch.from_numpy(_mem(p.W + 4 * zi * p.w_bs_inner, p.N * p.K).reshape(p.N, p.K).copy())
                acc = Am.double() @ W.double().T
                if p.bias:
                    acc = acc + torch.from_numpy(_mem(p.bias + 4 * zi * p.bias_bs_inner, p.N).copy()).double()
                if p.act == 1:
                    acc = 0.5 * acc * (1 + torch.erf(acc / math.sqrt(2.0)))
                elif p.act == 2:
                    acc = torch.relu(acc)
                elif glu:
                    acc = acc[:, 0::2] * torch.sigmoid(acc[:, 1::2])
                acc = (acc * p.alpha).float()
                c_off = zo * p.c_bs_outer + zi * p.c_bs_inner
                r_off = zo * p.r_bs_outer + zi * p.r_bs_inner
                m = torch.arange(p.M)
                seg, t = m // p.rows_per_seg, m % p.rows_per_seg
                orow = seg * p.out_rows_per_seg + t + p.out_row_off
                store = t < p.seg_rows_valid
                Cbuf = _mem(p.C + 4 * c_off, max_row * p.ldc)
                Cv = torch.from_numpy(Cbuf).as_strided((max_row, n_out), (p.ldc, 1))
                if p.residual:
                    Rv = torch.from_numpy(_mem(p.residual + 4 * r_off, max_row * p.ldr).copy()).as_strided(
                        (max_row, n_out), (p.ldr, 1))
                    acc = acc + Rv[orow]
                if p.seg_len:
                    nseg = zo * p.segs_per_outer + int(seg.max()) + 1
                    sl = torch.from_numpy(_mem(p.seg_len, nseg, np.int32).copy()).long()
                    acc[t >= sl[zo * p.segs_per_outer + seg]] = 0
                Cv[orow[store]] = acc[store]
        self.calls.append("gemm")
        return 0

    # ---- LayerNorm
    def cst_layernorm(self, x, ldx, gamma, beta, out_f32, out_lp, lp_dtype, ldo, rows, Cd, rows_per_seg,
                      seg_rows_valid, out_rows_per_seg, out_row_off, zero_invalid, stream):
        assert Cd in (512, 768)
        X = torch.from_numpy(_mem(x, rows * ldx).copy()).as_strided((rows, Cd), (ldx, 1))
        g = torch.from_numpy(_mem(gamma, Cd).copy())
        b = torch.from_numpy(_mem(beta, Cd).copy())
        Y = torch.nn.functional.layer_norm(X, (Cd,), g, b, 1e-5)
        m = torch.arange(rows)
        seg, t = m // rows_per_seg, m % rows_per_seg
        orow = seg * out_rows_per_seg + t + out_row_off
        valid = t < seg_rows_valid
        Y[~valid] = 0
        sel = torch.ones_like(valid) if zero_invalid else valid
        max_row = int(orow.max()) + 1
        for o in (out_f32, out_lp):
            if o:
                assert o is out_f32 or lp_dtype == F32
                Ov = torch.from_numpy(_mem(o, max_row * ldo)).as_strided((max_row, Cd), (ldo, 1))
                Ov[orow[sel]] = Y[sel]
        self.calls.append("layernorm")
        return 0

    def cst_posconv_pack(self, x, B, rows_per_seg, n_frames, xg, xg_dtype, t_pad_rows, stream):
        assert xg_dtype == F32
        X = torch.from_numpy(_mem(x, B * rows_per_seg * 768).copy()).view(B, rows_per_seg, 16, 48)
        G = torch.zeros(B, 16, t_pad_rows, 64)
        G[:, :, 64:64 + n_frames, :48] = X[:, :n_frames].permute(0, 2, 1, 3)
        _mem(xg, G.numel())[:] = G.reshape(-1).numpy()
        self.calls.append("posconv_pack")
        return 0

    def cst_broadcast_rows(self, src, rows, Cd, B, dst, stream):
        s = _mem(src, rows * Cd)
        _mem(dst, B * rows * Cd).reshape(B, rows * Cd)[:] = s[None, :]
        self.calls.append("broadcast_rows")
        return 0

    # ---- attention
    def cst_attention(self, q, k, v, out, dtype, ldq, ldkv, ldo, B, H, n_q, q_rps, n_kv, kv_rps, kv_len, stream):
        assert dtype == F32
        Q = torch.from_numpy(_mem(q, B * q_rps * ldq).copy()).as_strided((B, n_q, H, 64), (q_rps * ldq, ldq, 64, 1))
        K = torch.from_numpy(_mem(k, B * kv_rps * ldkv).copy()).as_strided((B, n_kv, H, 64), (kv_rps * ldkv, ldkv, 64, 1))
        V = torch.from_numpy(_mem(v, B * kv_rps * ldkv).copy()).as_strided((B, n_kv, H, 64), (kv_rps * ldkv, ldkv, 64, 1))
        s = torch.einsum("bqhd,bkhd->bhqk", Q.double(), K.double())
        if kv_len:
            kl = torch.from_numpy(_mem(kv_len, B, np.int32).copy()).long()
            s = s.masked_fill(torch.arange(n_kv)[None, None, None, :] >= kl[:, None, None, None], float("-inf"))
        o = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), V.double()).float()
        Ov = torch.from_numpy(_mem(out, B * q_rps * ldo)).as_strided((B, n_q, H, 64), (q_rps * ldo, ldo, 64, 1))
        Ov[:] = o
        self.calls.append("attention")
        return 0

    def cst_attention_segs(self, q, k, v, out, dtype, ldq, ldkv, ldo, B, H, seg, max_n_q, max_n_kv, q_total, kv_total, kv_len,
                           stream):
        assert dtype == F32
        tbl = _mem(seg, 4 * B, np.int32).reshape(B, 4)
        kl = _mem(kv_len, B, np.int32) if kv_len else None
        for b in range(B):
            q0, nq, k0, nk = (int(t) for t in tbl[b])
            assert nq <= max_n_q and nk <= max_n_kv and q0 + nq <= q_total and k0 + nk <= kv_total
            Q = torch.from_numpy(_mem(q + 4 * q0 * ldq, nq * ldq).copy()).as_strided((nq, H, 64), (ldq, 64, 1))
            K = torch.from_numpy(_mem(k + 4 * k0 * ldkv, nk * ldkv).copy()).as_strided((nk, H, 64), (ldkv, 64, 1))
            V = torch.from_numpy(_mem(v + 4 * k0 * ldkv, nk * ldkv).copy()).as_strided((nk, H, 64), (ldkv, 64, 1))
            s_ = torch.einsum("qhd,khd->hqk", Q.double(), K.double())
            if kl is not None:
                s_ = s_.masked_fill(torch.arange(nk)[None, None, :] >= int(kl[b]), float("-inf"))
            o = torch.einsum("hqk,khd->qhd", torch.softmax(s_, -1), V.double()).float()
            Ov = torch.from_numpy(_mem(out + 4 * q0 * ldo, nq * ldo)).as_strided((nq, H, 64), (ldo, 64, 1))
            Ov[:] = o
        self.calls.append("attention_segs")
        return 0


# ---- greedy decoding (cst_dec_*): same pointer-level semantics on host memory, fp32 -----------------------------------
def _dec_embed(self, tokens, ld_tok, embed, w_dtype, pos_table, scale, x, B, Cd, step, stream):
    assert w_dtype == F32
    t = int(_mem(step, 1, np.int32)[0])
    for b in range(B):
        tok = int(_mem(tokens + 4 * (b * ld_tok + t), 1, np.int32)[0])
        _mem(x + 4 * b * Cd, Cd)[:] = np.float32(scale) * _mem(embed + 4 * tok * Cd, Cd) + _mem(pos_table + 4 * t * Cd, Cd)
    self.calls.append("dec_embed")
    return 0


def _dec_linear(self, pref, stream):
    p = pref._obj
    assert p.a_dtype == F32 and p.w_dtype == F32 and p.K % 512 == 0 and 1 <= p.n_seg <= 3 and p.N % p.n_seg == 0
    assert all(p.out_dtype[s] == F32 for s in range(p.n_seg))
    step = int(_mem(p.step, 1, np.int32)[0]) if p.step else 0
    A = torch.from_numpy(np.stack([_mem(p.A + 4 * m * p.lda, p.K).copy() for m in range(p.M)]))
    W = torch.from_numpy(_mem(p.W, p.N * p.K).reshape(p.N, p.K).copy())
    if p.ln_gamma:
        assert p.K == 512
        A = torch.nn.functional.layer_norm(A, (p.K,), torch.from_numpy(_mem(p.ln_gamma, p.K).copy()),
                                           torch.from_numpy(_mem(p.ln_beta, p.K).copy()), 1e-5)
    acc = A.double() @ W.double().T
    if p.bias:
        acc = acc + torch.from_numpy(_mem(p.bias, p.N).copy()).double()
    if p.act == 2:
        acc = torch.relu(acc)
    else:
        assert p.act == 0
    acc = acc.float().numpy()
    if p.residual:
        assert p.n_seg == 1
        acc = acc + np.stack([_mem(p.residual + 4 * m * p.ldr, p.N).copy() for m in range(p.M)])
    sn = p.N // p.n_seg
    for s in range(p.n_seg):
        for m in range(p.M):
            _mem(p.out[s] + 4 * (m * p.ldo[s] + step * p.step_stride[s]), sn)[:] = acc[m, s * sn:(s + 1) * sn]
    self.calls.append("dec_linear")
    return 0


def _dec_attention(self, q, ldq, k, v, kv_dtype, kv_bs, kv_rs, out, ldo, B, H, n_keys, n_max, step, stream, kv_group=1):
    assert kv_dtype == F32
    n = min(int(_mem(step, 1, np.int32)[0]) + 1, n_max) if step else n_keys
    for b in range(B):
        s_ = b // kv_group
        qb = torch.from_numpy(_mem(q + 4 * b * ldq, H * 64).copy()).view(H, 64).double()
        K = torch.from_numpy(np.stack([_mem(k + 4 * (s_ * kv_bs + j * kv_rs), H * 64).copy() for j in range(n)])).view(n, H, 64).double()
        Vv = torch.from_numpy(np.stack([_mem(v + 4 * (s_ * kv_bs + j * kv_rs), H * 64).copy() for j in range(n)])).view(n, H, 64).double()
        s = torch.einsum("hd,nhd->hn", qb, K)
        o = torch.einsum("hn,nhd->hd", torch.softmax(s, -1), Vv)
        _mem(out + 4 * b * ldo, H * 64)[:] = o.reshape(-1).float().numpy()
    self.calls.append("dec_attention")
    return 0


def _dec_select(self, logits, V, B, tokens, ld_tok, pos_scores, ld_ps, done, out_len, counters, max_len, min_len, pad, eos,
                stream):
    cnt = _mem(counters, 3, np.int32)
    step = int(cnt[0])
    if step > max_len:
        return 0
    dn, ol = _mem(done, B, np.int32), _mem(out_len, B, np.int32)
    for b in range(B):
        lp = torch.log_softmax(torch.from_numpy(_mem(logits + 4 * b * V, V).copy()), -1)
        lp[pad] = -math.inf
        if step >= max_len:
            lp[:eos] = -math.inf
            lp[eos + 1:] = -math.inf
        elif step < min_len:
            lp[eos] = -math.inf
        nxt = int(lp.argmax())
        trow = _mem(tokens + 4 * b * ld_tok, ld_tok, np.int32)
        if not dn[b]:
            trow[step + 1] = nxt
            _mem(pos_scores + 4 * b * ld_ps, ld_ps)[step] = float(lp[nxt])
            if nxt == eos:
                dn[b], ol[b] = 1, step + 1
                cnt[2] += 1
        else:
            trow[step + 1] = eos
    cnt[0] = step + 1
    self.calls.append("dec_select")
    return 0


EmuLib.cst_dec_embed = _dec_embed
EmuLib.cst_dec_linear = _dec_linear
EmuLib.cst_dec_attention = _dec_attention


def _dec_attention_grouped(self, q, ldq, k, v, kv_dtype, kv_bs, kv_rs, out, ldo, B, H, n_keys, n_max, step, kv_group, stream):
    return _dec_attention(self, q, ldq, k, v, kv_dtype, kv_bs, kv_rs, out, ldo, B, H, n_keys, n_max, step, stream, kv_group=kv_group)


EmuLib.cst_dec_attention_grouped = _dec_attention_grouped
EmuLib.cst_dec_select = _dec_select


def _text_embed(self, tokens, lengths, embed, pos_table, scale, x, valid, B, T, rows_per_seg, Cd, V, stream):
    tok = _mem(tokens, B * T, np.int64).reshape(B, T)
    lens = _mem(lengths, B, np.int64)
    E = _mem(embed, V * Cd).reshape(V, Cd)
    pos = _mem(pos_table, (T + 2) * Cd).reshape(T + 2, Cd)
    out = _mem(x, B * rows_per_seg * Cd).reshape(B, rows_per_seg, Cd)
    out[:] = 0
    for b in range(B):
        for t in range(T):
            p = t + 2 if t < lens[b] else 1
            out[b, t] = np.float32(scale) * E[min(V - 1, max(0, int(tok[b, t])))] + pos[p]
    if valid:
        _mem(valid, B, np.int32)[:] = np.minimum(T, np.maximum(0, lens))
    self.calls.append("text_embed")
    return 0


EmuLib.cst_text_embed = _text_embed


# ---- beam search (experimental ABI): cst_dec_attention_beam / cst_dec_beam_select --------------------------------------
def _dec_attention_beam(self, q, ldq, k, v, kv_dtype, kv_bs, kv_rs, out, ldo, R, H, n_max, hist, ld_hist, step, stream):
    assert kv_dtype == F32
    cur = min(int(_mem(step, 1, np.int32)[0]), n_max - 1)
    n = cur + 1
    for r in range(R):
        rows = [r if j == cur else int(_mem(hist + 4 * (r * ld_hist + j), 1, np.int32)[0]) for j in range(n)]
        qb = torch.from_numpy(_mem(q + 4 * r * ldq, H * 64).copy()).view(H, 64).double()
        K = torch.from_numpy(np.stack([_mem(k + 4 * (rows[j] * kv_bs + j * kv_rs), H * 64).copy() for j in range(n)])).view(n, H, 64).double()
        Vv = torch.from_numpy(np.stack([_mem(v + 4 * (rows[j] * kv_bs + j * kv_rs), H * 64).copy() for j in range(n)])).view(n, H, 64).double()
        o = torch.einsum("hn,nhd->hd", torch.softmax(torch.einsum("hd,nhd->hn", qb, K), -1), Vv)
        _mem(out + 4 * r * ldo, H * 64)[:] = o.reshape(-1).float().numpy()
    self.calls.append("dec_attention_beam")
    return 0


def _dec_beam_select(self, pref, stream):
    p = pref._obj
    B, K, V, T, C = p.B, p.K, p.V, p.T, 2 * p.K
    cnt = _mem(p.counters, 3, np.int32)
    step = int(cnt[0])
    logits = _mem(p.logits, B * K * V).reshape(B, K, V)
    tin, tout = _mem(p.tok_in, B * K * T, np.int32).reshape(B, K, T), _mem(p.tok_out, B * K * T, np.int32).reshape(B, K, T)
    sin, sout = _mem(p.sc_in, B * K * T).reshape(B, K, T), _mem(p.sc_out, B * K * T).reshape(B, K, T)
    hin, hout = _mem(p.hist_in, B * K * T, np.int32).reshape(B, K, T), _mem(p.hist_out, B * K * T, np.int32).reshape(B, K, T)
    ign = _mem(p.ignore, B * K, np.int32).reshape(B, K)
    ft, fp = _mem(p.fin_tokens, B * K * T, np.int32).reshape(B, K, T), _mem(p.fin_pos, B * K * T).reshape(B, K, T)
    fs, fl = _mem(p.fin_score, B * K).reshape(B, K), _mem(p.fin_len, B * K, np.int32).reshape(B, K)
    nfin, finished = _mem(p.n_final, B, np.int32), _mem(p.finished, B, np.int32)
    for b in range(B):
        if step > p.max_len or finished[b]:
            continue
        nb = 1 if step == 0 else K
        lp = torch.log_softmax(torch.from_numpy(logits[b, :nb].copy()), -1)
        lp[lp != lp] = -math.inf
        lp[:, p.pad] = -math.inf
        if step >= p.max_len:
            lp[:, :p.eos] = -math.inf
            lp[:, p.eos + 1:] = -math.inf
        elif step < p.min_len:
            lp[:, p.eos] = -math.inf
        if step > 0:
            lp = lp + torch.from_numpy(sin[b, :, step - 1].copy()).unsqueeze(1)
        flat = lp.reshape(-1)
        ncand = min(C, flat.numel() - 1)
        # top candidates, ties by lowest flat index (stable sort of the negated values)
        order = torch.sort(-flat, stable=True).indices[:ncand]
        cv, cbeam, ctok = flat[order], (order // V).tolist(), (order % V).tolist()
        eos = [ctok[c] == p.eos and float(cv[c]) != -math.inf and not (c < K and ign[b, c]) for c in range(ncand)]
        nf = int(nfin[b])
        for c in range(min(K, ncand)):
            if not eos[c] or nf >= K:
                continue
            src = cbeam[c]
            cum = np.concatenate((sin[b, src, :step], [float(cv[c])])).astype(np.float32)
            ft[b, nf, :step] = tin[b, src, 1:step + 1]
            ft[b, nf, step] = p.eos
            fp[b, nf, :step + 1] = np.diff(np.concatenate(([np.float32(0)], cum))).astype(np.float32)
            fl[b, nf] = step + 1
            fs[b, nf] = np.float32(float(cv[c]) / (step + 1) ** p.len_penalty)
            nf += 1
        nfin[b] = nf
        if nf == K or step == p.max_len:
            finished[b] = 1
            cnt[2] += 1
            continue
        flag = [bool(ign[b, c]) or eos[c] if c < K else eos[c] for c in range(ncand)]
        active = [c for c in range(ncand) if not flag[c]][:K]
        n_clean = len(active)
        active += [c for c in range(ncand) if flag[c]][:K - n_clean]
        new_tok, new_sc, new_h = tin[b].copy(), sin[b].copy(), hin[b].copy()
        for i, c in enumerate(active):
            src = cbeam[c]
            ign[b, i] = int(i >= n_clean)
            tout[b, i, :step + 1] = new_tok[src, :step + 1]
            tout[b, i, step + 1] = ctok[c]
            sout[b, i, :step] = new_sc[src, :step]
            sout[b, i, step] = float(cv[c])
            hout[b, i, :step] = new_h[src, :step]
            hout[b, i, step] = b * K + src
    if step <= p.max_len:
        cnt[0] = step + 1
    self.calls.append("dec_beam_select")
    return 0


EmuLib.cst_dec_attention_beam = _dec_attention_beam
EmuLib.cst_dec_beam_select = _dec_beam_select


# ---- backward pass (cst_transpose .. cst_conv0_bwd): the derivative entry points of the training step (train.py), fp32, on host memory.
# Where a derivative is not a plain index shuffle it is obtained from torch autograd of the forward expression, so the emulator states
# WHAT the kernel computes independently of how csrc/backward.cu computes it. ------------------------------------------------------------
def _view(ptr, rows, cols, ld, copy=True):
    if rows <= 0:
        return torch.zeros(0, cols)
    t = torch.from_numpy(_mem(ptr, (rows - 1) * ld + cols))
    t = t.as_strided((rows, cols), (ld, 1))
    return t.clone() if copy else t


def _transpose(self, x, x_dtype, ldx, rows, cols, outT, out_dtype, rows_pad, chunk, copy, ldcopy, stream):
    assert x_dtype == F32 and out_dtype == F32 and rows_pad >= rows
    chunk = rows_pad if chunk <= 0 else chunk
    assert rows_pad % chunk == 0
    X = _view(x, rows, cols, ldx)                                  # ldx < cols: overlapping windows of a strided convolution's input
    Xp = torch.zeros(rows_pad, cols)
    Xp[:rows] = X
    out = Xp.view(rows_pad // chunk, chunk, cols).permute(0, 2, 1).contiguous()       # [(r / chunk), c, r % chunk]
    _mem(outT, out.numel())[:] = out.reshape(-1).numpy()
    if copy:
        _view(copy, rows, cols, ldcopy, copy=False)[:] = X
    self.calls.append("transpose")
    return 0


def _colsum(self, x, x_dtype, ldx, rows, cols, out, ws, scale, stream):
    assert x_dtype == F32
    _mem(out, cols)[:] = (_view(x, rows, cols, ldx).double().sum(0) * scale).float().numpy()
    self.calls.append("colsum")
    return 0


def _act(kind, z, alpha):
    if kind == 1:
        y = 0.5 * z * (1 + torch.erf(z / math.sqrt(2.0)))
    elif kind == 2:
        y = torch.relu(z)
    else:
        assert kind == 3
        y = z[:, 0::2] * torch.sigmoid(z[:, 1::2])               # interleaved (value, gate) pairs
    return y * alpha


def _act_fwd(self, act, z, z_dtype, ldz, rows, cols_out, y, y_dtype, ldy, alpha, stream):
    assert z_dtype == F32 and y_dtype == F32
    Z = _view(z, rows, cols_out * (2 if act == 3 else 1), ldz).double()
    _view(y, rows, cols_out, ldy, copy=False)[:] = _act(act, Z, alpha).float()
    self.calls.append("act_fwd")
    return 0


def _act_bwd(self, act, z, z_dtype, ldz, dy, dy_dtype, ldy, rows, cols_out, dz, dz_dtype, lddz, alpha, stream):
    assert z_dtype == F32 and dy_dtype == F32 and dz_dtype == F32
    cin = cols_out * (2 if act == 3 else 1)
    Z = _view(z, rows, cin, ldz).double().requires_grad_()
    _act(act, Z, alpha).backward(_view(dy, rows, cols_out, ldy).double())
    _view(dz, rows, cin, lddz, copy=False)[:] = Z.grad.float()
    self.calls.append("act_bwd")
    return 0


def _layernorm_bwd(self, x, ldx, gamma, dy, ldy, dx, lddx, part, rows, Cd, accumulate, stream):
    X = _view(x, rows, Cd, ldx).double().requires_grad_()
    g = torch.from_numpy(_mem(gamma, Cd).copy()).double()
    DY = _view(dy, rows, Cd, ldy).double()
    torch.nn.functional.layer_norm(X, (Cd,), g, None, 1e-5).backward(DY)
    out = _view(dx, rows, Cd, lddx, copy=False)
    out[:] = (out.double() + X.grad if accumulate else X.grad).float()
    if part:
        xd = X.detach()
        xhat = (xd - xd.mean(1, keepdim=True)) / torch.sqrt(xd.var(1, unbiased=False, keepdim=True) + 1e-5)
        nblk = (rows + 7) // 8
        both = torch.zeros(nblk * 8, 2 * Cd, dtype=torch.float64)
        both[:rows, :Cd], both[:rows, Cd:] = DY * xhat, DY
        _mem(part, nblk * 2 * Cd)[:] = both.view(nblk, 8, 2 * Cd).sum(1).float().reshape(-1).numpy()      # [block][dgamma | dbeta]
    self.calls.append("layernorm_bwd")
    return 0


def _attention_bwd(self, q, k, v, o, dtype, d_o, dq, dk, dv, ldq, ldkv, ldo_fwd, ldo, lddq, lddkv, B, H, n_q, q_rps, n_kv, kv_rps, kv_len,
                   stream):
    assert dtype == F32

    def heads(ptr, n, rps, ld, copy=True):
        t = torch.from_numpy(_mem(ptr, ((B - 1) * rps + n - 1) * ld + H * 64)).as_strided((B, n, H, 64), (rps * ld, ld, 64, 1))
        return t.clone() if copy else t
    Q, K, V = (heads(p_, n, r, ld).double().requires_grad_() for p_, n, r, ld in ((q, n_q, q_rps, ldq), (k, n_kv, kv_rps, ldkv),
                                                                                   (v, n_kv, kv_rps, ldkv)))
    s = torch.einsum("bqhd,bkhd->bhqk", Q, K)
    if kv_len:
        kl = torch.from_numpy(_mem(kv_len, B, np.int32).copy()).long()
        s = s.masked_fill(torch.arange(n_kv)[None, None, None, :] >= kl[:, None, None, None], float("-inf"))
    out = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), V)
    out.backward(heads(d_o, n_q, q_rps, ldo).double())
    heads(dq, n_q, q_rps, lddq, copy=False)[:] = Q.grad.float()
    for dst, grad in ((dk, K.grad), (dv, V.grad)):                # "must be zero on entry": the kernel accumulates with atomics
        view = heads(dst, n_kv, kv_rps, lddkv, copy=False)
        view[:] = view + grad.float()
    self.calls.append("attention_bwd")
    return 0


def _col2im(self, dcol, dcol_dtype, M, k, stride, Cd, dx, rows_in, accumulate, stream):
    assert dcol_dtype == F32
    D = torch.from_numpy(_mem(dcol, M * k * Cd).copy()).view(M, k, Cd).double()
    acc = torch.zeros(rows_in, Cd, dtype=torch.float64)
    for t in range(k):                                             # window m covers input rows m*stride .. m*stride + k - 1
        r = torch.arange(M) * stride + t
        ok = r < rows_in
        acc.index_add_(0, r[ok], D[ok, t])
    out = _view(dx, rows_in, Cd, Cd, copy=False)
    out[:] = (out.double() + acc if accumulate else acc).float()
    self.calls.append("col2im")
    return 0


def _rows_remap(self, inp, ldi, in_rps, in_off, out, out_dtype, ldo, out_rps, out_off, n_seg, n_rows, Cd, seg_valid, seg_len, accumulate,
                scale, stream):
    assert out_dtype == F32
    sl = _mem(seg_len, n_seg, np.int32) if seg_len else None
    for s in range(n_seg):
        valid = min(seg_valid, int(sl[s])) if sl is not None else seg_valid
        valid = max(0, min(valid, n_rows))
        val = torch.zeros(n_rows, Cd)
        if valid:
            val[:valid] = _view(inp + 4 * (s * in_rps + in_off) * ldi, valid, Cd, ldi) * scale
        dst = _view(out + 4 * (s * out_rps + out_off) * ldo, n_rows, Cd, ldo, copy=False)
        dst[:] = dst + val if accumulate else val
    self.calls.append("rows_remap")
    return 0


def _conv0_bwd(self, wave, B, L, w, gamma, beta, scale_shift, dout, rows_per_seg, dw, dgamma, dbeta, ws, grad_scale, stream):
    x = torch.from_numpy(_mem(wave, B * L).reshape(B, L).copy()).double()
    W = torch.from_numpy(_mem(w, 5120).reshape(512, 1, 10).copy()).double().requires_grad_()
    g = torch.from_numpy(_mem(gamma, 512).copy()).double().requires_grad_()
    b = torch.from_numpy(_mem(beta, 512).copy()).double().requires_grad_()
    y = torch.nn.functional.conv1d(x.unsqueeze(1), W, stride=5)
    y = torch.nn.functional.group_norm(y, 512, g, b, 1e-5)
    y = 0.5 * y * (1 + torch.erf(y / math.sqrt(2.0)))
    T0 = y.shape[2]
    D = torch.from_numpy(_mem(dout, B * rows_per_seg * 512).copy()).view(B, rows_per_seg, 512)[:, :T0].double()
    y.backward(D.transpose(1, 2))
    for dst, t in ((dw, W.grad), (dgamma, g.grad), (dbeta, b.grad)):
        _mem(dst, t.numel())[:] = (t * grad_scale).float().reshape(-1).numpy()      # GradMultiply folded into the three outputs
    self.calls.append("conv0_bwd")
    return 0


def _embed_bwd(self, tokens, dx, scale, dE, B, T, rows_per_seg, Cd, V, pad_idx, stream):
    tok = torch.from_numpy(_mem(tokens, B * T, np.int64).copy()).view(B, T)
    D = torch.from_numpy(_mem(dx, B * rows_per_seg * Cd).copy()).view(B, rows_per_seg, Cd)[:, :T]
    E = torch.from_numpy(_mem(dE, V * Cd)).view(V, Cd)
    keep = tok != pad_idx
    E.index_add_(0, tok[keep], D[keep] * scale)
    self.calls.append("embed_bwd")
    return 0


def _adam_step(self, p, g, g_dtype, m, v, n, lr, b1, b2, eps, wd, step_size, grad_scale, dyn, stream):
    assert g_dtype == F32
    if dyn:
        lr, step_size, grad_scale = (float(t) for t in _mem(dyn, 3))
    P, G, M_, V_ = (torch.from_numpy(_mem(t, n)) for t in (p, g, m, v))
    gi = G * grad_scale
    M_.mul_(b1).add_(gi, alpha=1 - b1)
    V_.mul_(b2).addcmul_(gi, gi, value=1 - b2)
    P.mul_(1 - wd * lr)
    P.addcdiv_(M_, V_.sqrt() + eps, value=-step_size)
    self.calls.append("adam_step")
    return 0


EmuLib.cst_transpose = _transpose
EmuLib.cst_colsum = _colsum
EmuLib.cst_act_fwd = _act_fwd
EmuLib.cst_act_bwd = _act_bwd
EmuLib.cst_layernorm_bwd = _layernorm_bwd
EmuLib.cst_attention_bwd = _attention_bwd
EmuLib.cst_col2im = _col2im
EmuLib.cst_rows_remap = _rows_remap
EmuLib.cst_conv0_bwd = _conv0_bwd
EmuLib.cst_embed_bwd = _embed_bwd
EmuLib.cst_adam_step = _adam_step


# ---- dropout: the numpy statement of csrc/philox.cuh (Philox4x32-10, key = seed, counter = {group lo, group hi, site, 0}) -------------
def philox4x32_10(seed, group, site):
    """group: uint64 array -> uint32 array [..., 4]"""
    group = np.asarray(group, dtype=np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    c0, c1 = group & m32, group >> np.uint64(32)
    c2, c3 = np.full_like(c0, np.uint64(site)), np.zeros_like(c0)
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(0xD2511F53) * c0, np.uint64(0xCD9E8D57) * c2
        c0, c1, c2, c3 = (p1 >> np.uint64(32)) ^ c1 ^ k0, p1 & m32, (p0 >> np.uint64(32)) ^ c3 ^ k1, p0 & m32
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & m32, (k1 + np.uint64(0xBB67AE85)) & m32
    return np.stack((c0, c1, c2, c3), -1).astype(np.uint32)


def dropout_keep(seed, site, n_elements, p):
    """keep mask (bool [n_elements]) of elements 0 .. n-1 of one dropout site; n_elements % 4 == 0"""
    assert n_elements % 4 == 0
    words = philox4x32_10(seed, np.arange(n_elements // 4, dtype=np.uint64), site).reshape(-1)
    return words >= np.uint32(min(int(float(np.float32(p)) * 4294967296.0), 0xFFFFFFFF))


def _dropout(self, x, x_dtype, ldx, add, ldadd, out, out_dtype, ldo, out2, out2_dtype, ldo2, rows, cols, p, seed, site, stream):
    assert x_dtype == F32 and out_dtype == F32 and (not out2 or out2_dtype == F32) and cols % 4 == 0
    sd = int(_mem(seed, 1, np.uint64)[0])
    keep = torch.from_numpy(dropout_keep(sd, site, rows * cols, p)).view(rows, cols)
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    val = torch.where(keep, _view(x, rows, cols, ldx) * float(scale), torch.zeros(()))
    if add:
        val = val + _view(add, rows, cols, ldadd)
    _view(out, rows, cols, ldo, copy=False)[:] = val
    if out2:
        _view(out2, rows, cols, ldo2, copy=False)[:] = val
    self.calls.append("dropout")
    self.dropout_log.append((sd, site, rows, cols, float(p)))
    return 0


EmuLib.cst_dropout = _dropout


def oracle_dropout_hook(step, pass_id, geom, p_of):
    """DROPOUT_HOOK for oracle/chimera_oracle.py: the masks an `EncoderTrainStep` drew, re-created from (seed, site) with the numpy
    Philox above and mapped from the kernel's padded row space [B * rows_per_seg, C] onto the oracle's [B, T, C] tensors.
    geom(tag) -> rows_per_seg of that site, p_of(tag) -> its probability.  -> (hook, list of tags it was called with)."""
    seed = int(step.seed_dev.item())
    used = []

    def hook(tag, x):
        p = p_of(tag)
        if p <= 0:
            return x
        site = step._sites[(pass_id, tag)]
        used.append(tag)
        scale = float(np.float32(1.0) / (np.float32(1.0) - np.float32(p)))
        if x.dim() == 4:                                            # attention probabilities [B, H, Tq, Tk]; geom -> (n_q, n_kv) of our call
            B, H, Tq, Tk = x.shape
            n_q, n_kv = geom(tag)
            keep = torch.from_numpy(dropout_keep(seed, site, B * H * _up64(n_q) * _up64(n_kv), p))
            return x * keep.view(B, H, _up64(n_q), _up64(n_kv))[:, :, :Tq, :Tk].to(x.dtype) * scale
        B, Tn, Cd = x.shape
        rps = geom(tag)
        keep = torch.from_numpy(dropout_keep(seed, site, B * rps * Cd, p)).view(B, rps, Cd)[:, :Tn]
        return x * keep.to(x.dtype) * scale
    return hook, used


# ---- attention with dropout of the probabilities (cst_attention_dropout_fwd / cst_attention_bwd_tc_dropout), fp32 on the host ----------
def _up64(n):
    return (n + 63) // 64 * 64


def _attn_prob_keep(seed_ptr, site, B, H, n_q, n_kv, p):
    """keep / (1 - p) factors [B, H, n_q, n_kv]: element ((b*H + h)*up64(n_q) + i)*up64(n_kv) + j of the site"""
    Tqp, Tkp = _up64(n_q), _up64(n_kv)
    if p <= 0:
        return torch.ones(B, H, n_q, n_kv, dtype=torch.float64)
    sd = int(_mem(seed_ptr, 1, np.uint64)[0])
    keep = torch.from_numpy(dropout_keep(sd, site, B * H * Tqp * Tkp, p)).view(B, H, Tqp, Tkp)[:, :, :n_q, :n_kv]
    return keep.double() * float(np.float32(1.0) / (np.float32(1.0) - np.float32(p)))


def _heads(ptr, B, H, n, rps, ld, copy=True):
    t = torch.from_numpy(_mem(ptr, ((B - 1) * rps + n - 1) * ld + H * 64)).as_strided((B, n, H, 64), (rps * ld, ld, 64, 1))
    return t.clone() if copy else t


def _attn_dropout_graph(self, q, k, v, ldq, ldkv, B, H, n_q, q_rps, n_kv, kv_rps, kv_len, p, seed, site):
    Q, K, V = (_heads(p_, B, H, n, r, ld).double().requires_grad_() for p_, n, r, ld in ((q, n_q, q_rps, ldq), (k, n_kv, kv_rps, ldkv),
                                                                                         (v, n_kv, kv_rps, ldkv)))
    s = torch.einsum("bqhd,bkhd->bhqk", Q, K)
    if kv_len:
        kl = torch.from_numpy(_mem(kv_len, B, np.int32).copy()).long()
        s = s.masked_fill(torch.arange(n_kv)[None, None, None, :] >= kl[:, None, None, None], float("-inf"))
    pd = torch.softmax(s, -1) * _attn_prob_keep(seed, site, B, H, n_q, n_kv, p)
    return Q, K, V, torch.einsum("bhqk,bkhd->bqhd", pd, V)


def _attention_dropout_fwd(self, q, k, v, qkv_dtype, out, out_dtype, ldq, ldkv, ldo, B, H, n_q, q_rps, n_kv, kv_rps, kv_len, p, seed, site,
                           ws, stream):
    assert qkv_dtype == F32 and out_dtype == F32
    _, _, _, o = _attn_dropout_graph(self, q, k, v, ldq, ldkv, B, H, n_q, q_rps, n_kv, kv_rps, kv_len, p, seed, site)
    _heads(out, B, H, n_q, q_rps, ldo, copy=False)[:] = o.detach().float()
    self.calls.append("attention_dropout_fwd")
    return 0


def _attention_bwd_tc_dropout(self, q, k, v, qkv_dtype, d_o, dq, dk, dv, ldq, ldkv, ldo, lddq, lddkv, B, H, n_q, q_rps, n_kv, kv_rps, kv_len,
                              p, seed, site, ws, stream):
    assert qkv_dtype == F32
    Q, K, V, o = _attn_dropout_graph(self, q, k, v, ldq, ldkv, B, H, n_q, q_rps, n_kv, kv_rps, kv_len, p, seed, site)
    o.backward(_heads(d_o, B, H, n_q, q_rps, ldo).double())
    _heads(dq, B, H, n_q, q_rps, lddq, copy=False)[:] = Q.grad.float()
    _heads(dk, B, H, n_kv, kv_rps, lddkv, copy=False)[:] = K.grad.float()            # written, not accumulated
    _heads(dv, B, H, n_kv, kv_rps, lddkv, copy=False)[:] = V.grad.float()
    self.calls.append("attention_bwd_tc_dropout")
    return 0


EmuLib.cst_attention_dropout_fwd_ws_bytes = lambda self, B, H, n_q, n_kv: 256
EmuLib.cst_attention_bwd_tc_ws_bytes = lambda self, B, H, n_q, n_kv: 256
EmuLib.cst_attention_dropout_fwd = _attention_dropout_fwd
EmuLib.cst_attention_bwd_tc_dropout = _attention_bwd_tc_dropout
