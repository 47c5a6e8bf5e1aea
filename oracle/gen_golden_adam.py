"""TEST INFRASTRUCTURE ONLY -- three steps of the UNMODIFIED reference optimizer (fairseq.optim.adam.Adam, fairseq/optim/adam.py)
on seeded tensors -> tests/golden/adam.npz.  Dev container only:  PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden_adam"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import make_overlay  # noqa: E402


def main():
    make_overlay.build()
    make_overlay.activate()
    import collections
    import collections.abc
    collections.Collection = collections.abc.Collection      # the reference imports the pre-3.10 alias (adam.py:8); environment shim only
    from fairseq.optim.adam import Adam
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(3, 257, generator=g)
    grads = [torch.randn(3, 257, generator=g) * (0.1 + i) for i in range(3)]
    hp = dict(lr=2e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.01)
    p = torch.nn.Parameter(p0.clone())
    opt = Adam([p], **hp)
    after = []
    for gr in grads:
        p.grad = gr.clone()
        opt.step()
        after.append(p.detach().clone().numpy())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "adam.npz"), p0=p0.numpy(), grads=np.stack([x.numpy() for x in grads]),
                        after=np.stack(after), lr=hp["lr"], betas=np.asarray(hp["betas"]), eps=hp["eps"], weight_decay=hp["weight_decay"])
    print("wrote adam.npz", after[-1].ravel()[:3])


if __name__ == "__main__":
    main()
