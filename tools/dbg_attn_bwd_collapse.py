"""attention_bwd precision when keys/values are nearly identical (rank-collapsed random-init activations)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import chimera_st_b200  # noqa
from chimera_st_b200.train import _Ops
o = _Ops(torch.device("cuda"))
g = torch.Generator().manual_seed(0)
B, H, Tq, Tk = 2, 8, 16, 5
Cd = H * 64
for spread in (1.0, 0.1, 0.01, 0.001):
    base = torch.randn(1, 1, Cd, generator=g)
    q0 = torch.randn(B, Tq, Cd, generator=g) * 0.4
    k0 = base + spread * torch.randn(B, Tk, Cd, generator=g)
    v0 = base * 2 + spread * torch.randn(B, Tk, Cd, generator=g)
    do = torch.randn(B, Tq, Cd, generator=g)
    def ref(dt):
        q, k, v = (t.to(dt).clone().requires_grad_() for t in (q0, k0, v0))
        qh, kh, vh = (t.view(B, -1, H, 64).transpose(1, 2) for t in (q, k, v))
        out = (torch.softmax(qh @ kh.transpose(-1, -2), -1) @ vh).transpose(1, 2).reshape(B, Tq, Cd)
        out.backward(do.to(dt))
        return out.detach(), q.grad, k.grad, v.grad
    o64, q64, k64, v64 = ref(torch.float64)
    o32, q32, k32, v32 = ref(torch.float32)
    qd, kd, vd = q0.cuda().view(B * Tq, Cd), k0.cuda().view(B * Tk, Cd), v0.cuda().view(B * Tk, Cd)
    ctx = o.attention(qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), Cd, Cd, B, H, Tq, Tq, Tk, Tk, None, B * Tq, Cd)
    dq, dk, dv = (torch.zeros_like(t) for t in (qd, kd, vd))
    o.attention_bwd(qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), ctx, do.cuda().view(B * Tq, Cd), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                    Cd, Cd, Cd, B, H, Tq, Tq, Tk, Tk, None)
    torch.cuda.synchronize()
    rl = lambda a, b: float((a.double().cpu().reshape(b.shape) - b).norm() / b.norm())
    print("spread %g: fwd ours %.2e torch32 %.2e | dq ours %.2e t32 %.2e | dk ours %.2e t32 %.2e | dv ours %.2e t32 %.2e" % (
        spread, rl(ctx, o64), rl(o32, o64), rl(dq, q64), rl(q32, q64), rl(dk, k64), rl(k32, k64), rl(dv, v64), rl(v32, v64)))
