"""GPU parity of the beam-search kernels (cst_dec_attention_beam, cst_dec_beam_select) and of B200BeamDecoder against the
UNMODIFIED reference generator's beam-5 hypotheses (tests/golden/beam.npz, oracle/gen_golden_beam.py)."""
import os

import numpy as np
import pytest
import torch

import chimera_st_b200  # noqa: F401
from chimera_st_b200 import synth, _lib as L
from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kvdt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("R,step", [(5, 0), (10, 7), (40, 101)])
def test_attention_with_history_table(R, step, kvdt):
    H, T = 8, 104
    g = torch.Generator().manual_seed(3)
    q = torch.randn(R, 512, generator=g) * 0.3
    K, V = torch.randn(R, T, 512, generator=g).to(kvdt), torch.randn(R, T, 512, generator=g).to(kvdt)
    hist = torch.randint(0, R, (R, T), generator=g, dtype=torch.int32)
    rows = hist.long().clone()
    rows[:, step] = torch.arange(R)
    n = step + 1
    Kg = torch.stack([K[rows[r, :n], torch.arange(n)] for r in range(R)]).double().view(R, n, H, 64)
    Vg = torch.stack([V[rows[r, :n], torch.arange(n)] for r in range(R)]).double().view(R, n, H, 64)
    s = torch.einsum("rhd,rnhd->rhn", q.double().view(R, H, 64), Kg)
    ref = torch.einsum("rhn,rnhd->rhd", torch.softmax(s, -1), Vg).reshape(R, 512)
    out = torch.empty(R, 512, device="cuda")
    st = torch.tensor([step], dtype=torch.int32, device="cuda")
    Kc, Vc, hc, qc = K.cuda(), V.cuda(), hist.cuda(), q.cuda()
    L.check(L.load().cst_dec_attention_beam(qc.data_ptr(), 512, Kc.data_ptr(), Vc.data_ptr(), L.DT[kvdt], T * 512, 512,
                                            out.data_ptr(), 512, R, H, T, hc.data_ptr(), T, st.data_ptr(), L.stream_ptr()))
    assert rel_l2(out.cpu(), ref) < 2e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", ["tiny", "c1mix"])
def test_beam_decoder_reproduces_reference_generator(name, dtype):
    from chimera_st_b200.decoder import B200BeamDecoder
    g = np.load(os.path.join(GOLDEN, "beam.npz"))
    dsd = synth.make_decoder_state_dict(seed=int(g["decoder_seed"]))
    E = dsd["decoder.embed_tokens.weight"].clone()
    E[2] *= float(g["eos_scale"])
    dsd["decoder.embed_tokens.weight"] = dsd["decoder.output_projection.weight"] = E
    mem = torch.from_numpy(g[name + "_memories"]).cuda().to(dtype)
    toks, sc = g[name + "_tokens"], g[name + "_scores"]
    for use_graph in (False, True):
        dec = B200BeamDecoder(dsd, beam=int(g["beam"]), dtype=dtype, device="cuda", use_graph=use_graph)
        hyps = dec.generate(mem, max_len=int(g["max_len_b"]))
        for b, hs in enumerate(hyps):
            if dtype == torch.float32:
                assert len(hs) == int((~np.isnan(sc[b])).sum())
                for k, h in enumerate(hs):
                    assert h["tokens"].tolist() == [x for x in toks[b, k].tolist() if x >= 0], (name, b, k)
                    assert abs(h["score"] - sc[b, k]) < 2e-4
            else:                                   # bf16: the best hypothesis must survive operand rounding
                assert hs[0]["tokens"].tolist() == [x for x in toks[b, 0].tolist() if x >= 0] or abs(hs[0]["score"] - sc[b, 0]) < 0.1
