"""TEST INFRASTRUCTURE ONLY -- builds the import overlay that lets the UNMODIFIED
reference (/root/reference, a Sept-2020 fairseq fork) run under Python 3.12 /
torch 2.11 in the dev container.  Nothing is copied from the reference except
one file that is regex-patched on the fly into the git-ignored `oracle/_ref/`
(fairseq/dataclass/configs.py: py3.12 rejects its mutable dataclass defaults,
configs.py:880-889).  All arithmetic files (wav2vec2.py, multihead_attention.py,
transformer_layer.py, s2t_transformer.py, chimera/*.py) execute unmodified.

Where the reference comes from (`ref_root()`): /root/reference when it is mounted (dev
container), else the bundle `oracle/_ref/src/` that `oracle/make_ref_bundle.py` wrote -- verbatim
copies of the reference files this path imports, git-ignored, shipped to the GPU box with the
snapshot like a built .so.  The overlay resolves the reference root when it is IMPORTED (no absolute
path is baked in), so the same `oracle/_ref/` works here and on the box.  CST_REF_ROOT overrides.

usage: python oracle/make_overlay.py   ->  oracle/_ref/{ovl,shim}
"""
import os
import re
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
BUNDLE = os.path.join(OUT, "src")


def ref_root():
    """Root of the reference tree to import: CST_REF_ROOT, the mounted tree, or the bundle; None if there is none."""
    env = os.environ.get("CST_REF_ROOT")
    for cand in ([env] if env else [REF, BUNDLE]):
        if cand and os.path.isdir(os.path.join(cand, "fairseq")):
            return cand
    return None


def available():
    return ref_root() is not None


# evaluated inside the overlay package when it is imported
_RESOLVE = (
    "import os, sys\n"
    "def _ref_root():\n"
    "    here = os.path.dirname(os.path.abspath(__file__))\n"
    "    while os.path.basename(here) != 'ovl':\n"
    "        here = os.path.dirname(here)\n"
    "    env = os.environ.get('CST_REF_ROOT')\n"
    "    for cand in ([env] if env else ['/root/reference', os.path.join(os.path.dirname(here), 'src')]):\n"
    "        if cand and os.path.isdir(os.path.join(cand, 'fairseq')):\n"
    "            return cand\n"
    "    raise ImportError('no reference tree: neither /root/reference nor oracle/_ref/src')\n")


def _w(path, text):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(text)


def build(out=OUT, ref=None):
    ref = ref or ref_root()
    if ref is None:
        raise RuntimeError("no reference tree: neither %s nor the bundle %s (run oracle/make_ref_bundle.py in the "
                           "dev container)" % (REF, BUNDLE))
    shim, ovl = os.path.join(out, "shim"), os.path.join(out, "ovl")
    # --- stubs for packages the container lacks -------------------------------
    _w(os.path.join(shim, "omegaconf", "__init__.py"),
       "import contextlib\n"
       "def II(x): return '${%s}' % x\n"
       "MISSING = '???'\n"
       "class DictConfig(dict): pass\n"
       "class OmegaConf:\n"
       "    create = staticmethod(lambda x=None: DictConfig(x or {}))\n"
       "    set_struct = staticmethod(lambda *a, **k: None)\n"
       "    to_container = staticmethod(lambda x, **k: dict(x))\n"
       "@contextlib.contextmanager\n"
       "def open_dict(x): yield x\n"
       "class _utils: pass\n")
    _w(os.path.join(shim, "hydra", "__init__.py"), "")
    _w(os.path.join(shim, "hydra", "core", "__init__.py"), "")
    _w(os.path.join(shim, "hydra", "core", "config_store.py"),
       "class ConfigStore:\n"
       "    _i = None\n"
       "    @classmethod\n"
       "    def instance(cls):\n"
       "        cls._i = cls._i or cls(); return cls._i\n"
       "    def store(self, *a, **k): pass\n")
    _w(os.path.join(shim, "hydra", "core", "global_hydra.py"),
       "class GlobalHydra:\n"
       "    _i = None\n"
       "    @classmethod\n"
       "    def instance(cls):\n"
       "        cls._i = cls._i or cls(); return cls._i\n"
       "    def is_initialized(self): return False\n"
       "    def clear(self): pass\n")
    _w(os.path.join(shim, "hydra", "experimental", "__init__.py"),
       "def compose(*a, **k): raise NotImplementedError\n"
       "def initialize(*a, **k): raise NotImplementedError\n")
    for m in ("sacrebleu", "editdistance", "soundfile", "bitarray"):
        _w(os.path.join(shim, m, "__init__.py"), "")
    # --- overlay package: search overlay dir first, then the reference ---------
    _w(os.path.join(ovl, "fairseq", "__init__.py"),
       _RESOLVE +
       "__path__ = [os.path.dirname(__file__), os.path.join(_ref_root(), 'fairseq')]\n"
       "__version__ = '1.0.0a0'\n"
       "from fairseq.logging import meters, metrics, progress_bar\n"
       "sys.modules['fairseq.meters'] = meters\n"
       "sys.modules['fairseq.metrics'] = metrics\n"
       "sys.modules['fairseq.progress_bar'] = progress_bar\n")
    _w(os.path.join(ovl, "fairseq", "dataclass", "__init__.py"),
       _RESOLVE +
       "__path__ = [os.path.dirname(__file__), os.path.join(_ref_root(), 'fairseq', 'dataclass')]\n"
       "from .configs import FairseqDataclass\n"
       "from .constants import ChoiceEnum\n")
    src = open(os.path.join(ref, "fairseq", "dataclass", "configs.py")).read()
    src = re.sub(r"^(    [a-z_]+: ([A-Za-z]+Config)) = \2\(\)$",
                 r"\1 = field(default_factory=\2)", src, flags=re.M)
    _w(os.path.join(ovl, "fairseq", "dataclass", "configs.py"), src)
    return ovl, shim


def activate(out=OUT):
    """Put the overlay on sys.path and apply the numpy/torch compat preamble."""
    import numpy as np
    os.environ["TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD"] = "1"
    sys.dont_write_bytecode = True            # the reference tree is read-only
    for n, t in (("float", float), ("int", int), ("bool", bool),
                 ("object", object), ("str", str), ("complex", complex)):
        if n not in np.__dict__:
            setattr(np, n, t)
    import torch.nn.functional as F
    if not getattr(F.multi_head_attention_forward, "_cst_wrapped", False):
        _orig = F.multi_head_attention_forward

        def _mha(*a, **k):
            # torch>=2 baddbmm rejects the reference's float64 attn_mask
            # (w2v2_transformer_interlingua.py:284-287); torch<=1.7 accumulated it
            # in the query dtype, which this cast reproduces.
            a = list(a)
            if len(a) > 16 and a[16] is not None and a[16].is_floating_point():
                a[16] = a[16].to(a[0].dtype)
            return _orig(*a, **k)
        _mha._cst_wrapped = True
        F.multi_head_attention_forward = _mha
    for p in (os.path.join(out, "shim"), os.path.join(out, "ovl")):
        if p not in sys.path:
            sys.path.insert(0, p)


if __name__ == "__main__":
    print(build())
