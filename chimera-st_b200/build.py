"""Builds the C-ABI CUDA library in-tree (sm_100a only; nvcc cross-compiles without a GPU).

    python chimera-st_b200/build.py [--force]

Output: chimera-st_b200/libchimera_st_b200.so (git-ignored; travels to the GPU box with the snapshot).
"""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libchimera_st_b200.so")
SOURCES = ["api.cu", "lengths.cu", "wave.cu", "conv0.cu", "conv0_tc.cu", "norm.cu", "split.cu", "gemm_f32.cu", "gemm_tc.cu", "attention.cu", "attention_tc.cu", "attention_tc2.cu", "posconv_tc.cu", "posconv_tc2.cu", "decoder.cu", "text.cu", "loss.cu", "backward.cu", "attention_bwd.cu", "attention_bwd_tc.cu", "conv0_bwd.cu", "decoder_beam.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-DCST_BUILD"] + os.environ.get("CST_EXTRA_NVCC_FLAGS", "").split()


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = _nvcc()

    def cc(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj, r.stderr

    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(cc, SOURCES))
    if verbose:
        for _, err in res:
            print(err)
    cmd = [nvcc, "-shared", "-o", LIB] + [o for o, _ in res] + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    open(stamp, "w").write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
