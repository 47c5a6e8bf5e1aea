"""TEST INFRASTRUCTURE: compiles the reference's own Cython batching kernel
(/root/reference/fairseq/data/data_utils_fast.pyx) from where it lies into oracle/_ref/cy/ (git-ignored),
to pin `chimera_st_b200.batching.batch_by_size`.  Dev container only; the golden it produces
(tests/golden/batches.npz) is what travels."""
import os
import subprocess
import sys
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "cy")
SRC = "/root/reference/fairseq/data/data_utils_fast.pyx"


def build():
    os.makedirs(OUT, exist_ok=True)
    cpp = os.path.join(OUT, "data_utils_fast.cpp")
    subprocess.check_call([sys.executable, "-m", "cython", "--cplus", "-3", SRC, "-o", cpp])
    so = os.path.join(OUT, "data_utils_fast" + sysconfig.get_config_var("EXT_SUFFIX"))
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-w", "-I", sysconfig.get_paths()["include"],
                           "-I", np.get_include(), cpp, "-o", so])
    return so


if __name__ == "__main__":
    print(build())
    sys.path.insert(0, OUT)
    sys.path.insert(0, os.path.dirname(HERE))
    import data_utils_fast as ref
    import chimera_st_b200  # noqa: F401
    from chimera_st_b200 import batching
    rng = np.random.RandomState(2024)
    cases = {}
    for name, n, lo, hi, mt, ms, mult in (("c3", 512, 32000, 480000, 2000000, 0, 8), ("small", 57, 100, 3000, 9000, 0, 8),
                                          ("maxsent", 100, 500, 5000, 40000, 12, 4), ("mult1", 64, 1000, 20000, 50000, 0, 1)):
        lens = rng.randint(lo, hi + 1, size=n).astype(np.int64)
        idx = np.asarray(batching.ordered_indices(lens), dtype=np.int64)
        got = ref.batch_by_size_fast(idx, lambda i: int(lens[i]), mt, ms, mult)
        flat = np.concatenate([np.asarray(b, dtype=np.int64) for b in got])
        sizes = np.asarray([len(b) for b in got], dtype=np.int64)
        cases[name + "_lens"], cases[name + "_flat"], cases[name + "_sizes"] = lens, flat, sizes
        cases[name + "_cfg"] = np.asarray([mt, ms, mult], dtype=np.int64)
        mine = batching.batch_by_size(idx.tolist(), lens, mt, ms, mult)
        assert [list(b) for b in got] == mine, name
        print(name, len(got), "batches; restatement identical")
    np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "batches.npz"), **cases)
