"""TEST INFRASTRUCTURE ONLY -- the UNMODIFIED reference learning-rate schedule of the training recipe (--lr-scheduler inverse_sqrt --lr 1e-4
--warmup-updates 4000, chimera/scripts/train-en2any-ST.sh:48-49; fairseq/optim/lr_scheduler/inverse_square_root_schedule.py) sampled at a
set of update numbers -> tests/golden/lr_schedule.npz.  Dev container only:  PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden_lr"""
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import make_overlay  # noqa: E402


class _Opt:                                                   # the two methods the scheduler calls on its optimizer
    def __init__(self):
        self.lr = None

    def set_lr(self, lr):
        self.lr = lr

    def get_lr(self):
        return self.lr


def main():
    make_overlay.build()
    make_overlay.activate()
    import collections
    import collections.abc
    collections.Collection = collections.abc.Collection      # pre-3.10 alias the reference imports; environment shim only
    from fairseq.optim.lr_scheduler.inverse_square_root_schedule import InverseSquareRootSchedule
    from fairseq.optim.lr_scheduler.fairseq_lr_scheduler import FairseqLRScheduler
    out = {}
    updates = np.asarray([0, 1, 2, 10, 1999, 3999, 4000, 4001, 10000, 150000], dtype=np.int64)
    for name, (lr, warm, init) in {"recipe": (1e-4, 4000, -1.0), "init": (2e-3, 10, 1e-7)}.items():
        cfg = SimpleNamespace(lr=[lr], warmup_updates=warm, warmup_init_lr=init)
        opt = _Opt()
        # FairseqLRScheduler.__init__ insists on a FairseqOptimizer instance; the schedule itself only uses set_lr / get_lr
        orig = FairseqLRScheduler.__init__
        FairseqLRScheduler.__init__ = lambda self, cfg_, optimizer: (setattr(self, "cfg", cfg_), setattr(self, "optimizer", optimizer),
                                                                     setattr(self, "best", None)) and None
        try:
            sch = InverseSquareRootSchedule(cfg, opt)
        finally:
            FairseqLRScheduler.__init__ = orig
        first = opt.lr
        vals = [sch.step_update(int(u)) for u in updates]
        out[name + "_cfg"] = np.asarray([lr, warm, init])
        out[name + "_initial"] = np.float64(first)
        out[name + "_lr"] = np.asarray(vals, dtype=np.float64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lr_schedule.npz"), updates=updates, **out)
    print("wrote lr_schedule.npz", out["recipe_lr"])


if __name__ == "__main__":
    main()
