"""TEST / BASELINE INFRASTRUCTURE ONLY -- makes the UNMODIFIED reference travel to the GPU box.

The reference (Glaciohound/Chimera-ST) is pure Python on this path and `/root/reference` does not exist on the
GPU box, so `bench.py --impl reference`, `cpu_baseline` and the `-m gpu` plugin test could otherwise only time /
call the oracle *restatement*.  This script exercises the reference in the dev container (model construction
through the registry, encoder forward, the reference's SequenceGenerator, the --user-dir import hook, the
collater helpers), records every module that was imported from `/root/reference`, and copies exactly those
files VERBATIM to `oracle/_ref/src/` (git-ignored -- never part of the history -- but not gpurun-ignored, so
it ships with the snapshot like a built .so).  `oracle/make_overlay.py` then resolves the reference root to
that bundle when `/root/reference` is absent.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_ref_bundle            # discover + copy + self-check

The self-check re-runs the exercise in a subprocess with CST_REF_ROOT pointing at the bundle and fails if any
module is still loaded from /root/reference or if the memories differ from the mounted tree's.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import make_overlay  # noqa: E402

REF = make_overlay.REF
BUNDLE = make_overlay.BUNDLE
# files that are read, not imported
EXTRA = ["fairseq/dataclass/configs.py",                                   # source of the patched overlay copy
         "fairseq/version.txt",
         "chimera/resources/wmt14-en-de-spm/spm_unigram10000_wave_joint.txt",   # target dictionary (V = 10 000)
         "LICENSE"]


def exercise(plugin=True):
    """Run every reference code path the repo's tests / bench use; returns a checksum of the memories."""
    import argparse
    import torch
    make_overlay.build()
    make_overlay.activate()
    import chimera_st_b200  # noqa: F401
    from chimera_st_b200 import synth
    from oracle.gen_golden_greedy import build_reference_model
    model, d = build_reference_model()
    from fairseq.sequence_generator import SequenceGenerator
    import fairseq.search  # noqa: F401
    from fairseq import utils, checkpoint_utils, options, tasks  # noqa: F401
    from fairseq.tasks.fairseq_task import FairseqTask  # noqa: F401
    from fairseq.data import data_utils  # noqa: F401
    from fairseq.data.audio import speech_to_text_dataset, audio_utils, triplet_dataset  # noqa: F401
    import fairseq.criterions.triplet_st_mt_contrastive  # noqa: F401   (training heads)
    import fairseq.criterions.label_smoothed_cross_entropy  # noqa: F401
    import fairseq.legacy_distributed_data_parallel  # noqa: F401          (C5 all-reduce)
    wave, lens = synth.make_waveforms([9000, 6000], seed=5)
    sample = {"net_input": {"src_tokens": wave, "src_lengths": lens}}
    with torch.no_grad():
        mem = model.encoder(wave, lens).encoder_out
        for beam in (1, 5):
            SequenceGenerator([model], d, beam_size=beam, max_len_a=0, max_len_b=4).generate([model], sample)
    if plugin:
        utils.import_user_module(argparse.Namespace(user_dir=os.path.join(ROOT, "chimera-st_b200", "fairseq_plugin")))
    return hashlib.sha256(mem.numpy().tobytes()).hexdigest()


def imported_reference_files(root):
    out = set()
    for m in list(sys.modules.values()):
        f = getattr(m, "__file__", None)
        if f and os.path.abspath(f).startswith(root + os.sep):
            out.add(os.path.relpath(os.path.abspath(f), root))
    return out


def main():
    if not os.path.isdir(REF):
        raise SystemExit("needs the mounted reference tree %s (dev container)" % REF)
    os.environ.pop("CST_REF_ROOT", None)
    if "--check" in sys.argv:                     # subprocess mode: run from the bundle, report what was imported
        os.environ["CST_REF_ROOT"] = BUNDLE
        digest = exercise()
        leaked = sorted(imported_reference_files(REF))
        print(json.dumps({"digest": digest, "leaked": leaked, "from_bundle": len(imported_reference_files(BUNDLE))}))
        return
    digest = exercise()
    files = sorted(imported_reference_files(REF) | set(EXTRA))
    if os.path.isdir(BUNDLE):
        shutil.rmtree(BUNDLE)
    nbytes = 0
    for rel in files:
        src, dst = os.path.join(REF, rel), os.path.join(BUNDLE, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)                  # verbatim
        nbytes += os.path.getsize(dst)
    # namespace directories the reference auto-imports with os.listdir need their __init__.py (already in `files`)
    with open(os.path.join(BUNDLE, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": files, "bytes": nbytes,
                   "note": "verbatim copies of the reference files imported on the speech-encoding path; git-ignored"}, f, indent=1)
    print("bundle: %d files, %.2f MB -> %s" % (len(files), nbytes / 1e6, BUNDLE))
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-m", "oracle.make_ref_bundle", "--check"], cwd=ROOT, env=env,
                       capture_output=True, text=True)
    last = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if r.returncode != 0 or not last:
        raise SystemExit("bundle self-check failed:\n" + r.stdout[-2000:] + r.stderr[-4000:])
    res = json.loads(last[-1])
    assert not res["leaked"], "still imported from the mounted tree: %s" % res["leaked"][:10]
    assert res["digest"] == digest, "memories from the bundle differ from the mounted reference"
    print("self-check ok: %d modules from the bundle, memories bit-identical" % res["from_bundle"])


def build():
    """Called by __graft_entry__.build() when /root/reference is mounted; cheap if the bundle is current."""
    man = os.path.join(BUNDLE, "MANIFEST.json")
    if os.path.exists(man):
        try:
            files = json.load(open(man))["files"]
            if all((not os.path.exists(os.path.join(REF, f))) or
                   os.path.getsize(os.path.join(REF, f)) == os.path.getsize(os.path.join(BUNDLE, f)) for f in files):
                return BUNDLE
        except Exception:
            pass
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    env.pop("CST_REF_ROOT", None)
    subprocess.check_call([sys.executable, "-m", "oracle.make_ref_bundle"], cwd=ROOT, env=env)
    return BUNDLE


if __name__ == "__main__":
    main()
