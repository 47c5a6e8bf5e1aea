"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's Adam step (fairseq/optim/adam.py:157-224, the fp32 path that
FP16Optimizer drives) in plain torch tensor arithmetic; pinned to the unmodified class by tests/golden/adam.npz
(oracle/gen_golden_adam.py).  The product never imports this file."""
import math


def adam_step(p, g, m, v, t, lr, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.0):
    """In-place update of (p, m, v) for step number t (1-based), gradient g (fp32)."""
    b1, b2 = betas
    m.mul_(b1).add_(g, alpha=1 - b1)                          # :204
    v.mul_(b2).addcmul_(g, g, value=1 - b2)                   # :205
    denom = v.sqrt().add_(eps)                                # :212
    step_size = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)   # :214-216
    if weight_decay != 0:
        p.add_(p, alpha=-weight_decay * lr)                   # :218-221
    p.addcdiv_(m, denom, value=-step_size)                    # :223
    return p
