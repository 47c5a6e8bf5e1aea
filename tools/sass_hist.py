"""Opcode histograms of the shipped library's SASS, one file per kernel family -> profiles/sass_<round>_<family>.txt

    python tools/sass_hist.py r02

Evidence that the hot kernels are tcgen05 / TMEM / TMA code: UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), LDTM / STTM
(tcgen05.ld / st), UTMALDG / UTMASTG (cp.async.bulk.tensor load / store), UTCBAR (tcgen05.commit), SYNCS (mbarrier).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "chimera-st_b200", "libchimera_st_b200.so")
FAMILIES = {"gemm_tc": r"gemm_tc", "attention_tc": r"attention_tc", "conv0_tc": r"conv0_tc", "posconv": r"posconv",
            "layernorm": r"layernorm", "decoder": r"dec_", "train": r"(_bwd|_grad|loss|adam|dropout|colsum|head_pack|transpose)"}
KEY = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCCP", "SYNCS", "MUFU", "HMMA", "FFMA2", "FFMA")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            kernels[name][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    pretty = dict(zip(kernels, demangle))
    for fam, pat in FAMILIES.items():
        sel = [(k, c) for k, c in kernels.items() if re.search(pat, pretty[k])]
        if not sel:
            continue
        path = os.path.join(ROOT, "profiles", "sass_%s_%s.txt" % (tag, fam))
        with open(path, "w") as f:
            f.write("# cuobjdump -sass %s : opcode histograms, kernels matching /%s/ (sm_100a)\n" % (os.path.basename(LIB), pat))
            tot = collections.Counter()
            for k, c in sel:
                base = collections.Counter()
                for op, n in c.items():
                    base[op.split(".")[0]] += n
                tot.update(c)
                f.write("\n== %s\n   instructions: %d\n" % (pretty[k][:200], sum(c.values())))
                f.write("   key opcodes: " + ", ".join("%s=%d" % (o, base[o]) for o in KEY if base[o]) + "\n")
                variants = [(op, n) for op, n in sorted(c.items()) if op.split(".")[0] in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR")]
                f.write("   tensor/TMA variants: " + ", ".join("%s=%d" % v for v in variants) + "\n")
                f.write("   top 12: " + ", ".join("%s=%d" % v for v in c.most_common(12)) + "\n")
            base = collections.Counter()
            for op, n in tot.items():
                base[op.split(".")[0]] += n
            f.write("\n== family total: " + ", ".join("%s=%d" % (o, base[o]) for o in KEY if base[o]) + "\n")
        print(path, len(sel), "kernels")


if __name__ == "__main__":
    main()
