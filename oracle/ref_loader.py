"""Load the UNMODIFIED reference (/root/reference) on CPU as ground truth.

TEST INFRASTRUCTURE ONLY.  This module is used by `oracle/make_golden.py` (to
generate the committed fixtures under tests/golden/), by `tests/test_reference_dropin.py`
/ `tests/test_full_shape.py` (the reference's own code as the checker, skipped when neither
/root/reference nor the staged copy oracle/_ref exists) and by `bench.py --impl reference[-gpu]`.
Nothing in `waldo_b200/` imports it.

Recipe = SURVEY.md Appendix A: three stubs (matplotlib, models.modules.mat,
lpips), `Tensor.cuda` identity on CPU (wif.py:31), options resolved by the
reference's own parser from its own launch scripts.
"""
from __future__ import annotations

import contextlib
import os
import re
import shlex
import sys
import tempfile
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    """/root/reference in the build container; on the GPU box the byte-for-byte staged copy oracle/_ref
    (oracle/build_ref.py)."""
    for cand in (os.environ.get("WALDO_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "models", "nets", "lvd.py")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "nets", "lvd.py"))


def is_staged_copy() -> bool:
    return os.path.abspath(REF_ROOT) == os.path.join(_HERE, "_ref")


_loaded = {}


def _install_stubs():
    stub_dir = tempfile.mkdtemp(prefix="waldo_ref_stubs_")
    mpl = os.path.join(stub_dir, "matplotlib")
    os.makedirs(mpl)
    open(os.path.join(mpl, "__init__.py"), "w").close()
    with open(os.path.join(mpl, "colors.py"), "w") as f:
        f.write("class ListedColormap:\n    def __init__(self, *a, **k):\n        pass\n")
    with open(os.path.join(mpl, "path.py"), "w") as f:
        f.write("class Path:\n    def __init__(self, *a, **k):\n        pass\n")
    sys.path[:0] = [stub_dir, REF_ROOT]
    m = types.ModuleType("models.modules.mat")
    m.MatInpainter = type("MatInpainter", (torch.nn.Module,), {})
    sys.modules["models.modules.mat"] = m
    l = types.ModuleType("lpips")
    l.LPIPS = object
    sys.modules["lpips"] = l


def load():
    """Returns a namespace with the reference's modules (imported once)."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _install_stubs()
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # wif.py:31
    # tools/ has no __init__.py in the reference either: a namespace package rooted at REF_ROOT
    ns = types.SimpleNamespace()
    import tools.utils as ref_utils  # noqa
    import models.modules.warp as ref_warp  # noqa
    import models.nets.lvd as ref_lvd  # noqa
    import models.nets.wif as ref_wif  # noqa
    ns.utils, ns.warp, ns.lvd, ns.wif = ref_utils, ref_warp, ref_lvd, ref_wif
    _loaded["ns"] = ns
    return ns


def parse_opts(script: str = "scripts/cityscapes/test.sh", extra: str = ""):
    """Resolve the `synthesizer` option namespace exactly as the reference does
    for one of its launch scripts (tools/options.py:772-801)."""
    load()
    from tools.options import Options

    txt = open(os.path.join(REF_ROOT, script)).read()
    txt = txt.replace("\\\n", " ")
    m = re.search(r"helpers/synthesizer_\w+\.py(.*)", txt, flags=re.S)
    args = m.group(1)
    args = re.sub(r'"checkpoints/"\$\{\w+\}', "none", args)
    args = re.sub(r"\$\{DATETIME\}", "dt", args)
    args = re.sub(r"\$\{GPU_IDS\}", "0", args)
    args = re.sub(r"\$\{\w+\}", "x", args)
    argv = ["x"] + shlex.split(args) + shlex.split(extra)
    os.environ.setdefault("LOCAL_RANK", "0")
    old = sys.argv
    sys.argv = argv
    try:
        with contextlib.redirect_stdout(open(os.devnull, "w")):
            opt = Options().parse(load_synthesizer=True)["synthesizer"]
    finally:
        sys.argv = old
    return opt


@contextlib.contextmanager
def stable_sort():
    """Force `Tensor.sort(stable=True)` (lowest source index wins) while the
    reference's InverseWarp runs (warp.py:114) -- SURVEY.md §8c tie rule."""
    orig = torch.Tensor.sort

    def patched(self, *a, **k):
        k["stable"] = True
        if a:
            k["dim"] = a[0]
            a = ()
        return orig(self, **k)

    torch.Tensor.sort = patched
    try:
        yield
    finally:
        torch.Tensor.sort = orig
