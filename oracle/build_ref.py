"""Recipe that makes the UNMODIFIED reference travel to the GPU box: oracle/_ref/.

TEST INFRASTRUCTURE.  The reference's hot path is pure Python (no compiled code), so "building" it means staging the
files its import chain needs, byte for byte, under oracle/_ref/ (git-ignored, NOT gpurun-ignored: like a built .so it
rides along with the snapshot).  Nothing is copied into the tracked tree.  `__graft_entry__.build()` runs this in the
build container, where /root/reference exists; on the GPU box only the staged copy is read.

Staged: models/ (without the vendored MAT inpainter, which oracle/ref_loader.py stubs), tools/utils.py,
tools/options.py, scripts/ (the launch scripts are the reference's de-facto config files, resolved by its own parser).
Users: oracle/ref_loader.py -> tests (parity against the reference's own code at the benchmarked shapes) and
bench.py --impl reference / reference-gpu (the comparator arms).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("WALDO_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")

FILES = [
    "models/__init__.py",
    "models/nets/__init__.py", "models/nets/lvd.py", "models/nets/wif.py", "models/nets/flp.py",
    "models/modules/__init__.py", "models/modules/warp.py", "models/modules/conv.py", "models/modules/transform.py",
    "models/modules/edge.py", "models/modules/gan_loss.py", "models/modules/perceptual.py", "models/modules/spectral.py",
    "models/modules/weight_init.py",
    "tools/utils.py", "tools/options.py",
]
DIRS = ["scripts"]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def build(verbose: bool = False) -> str | None:
    """Stage the reference under oracle/_ref.  Returns the path, or None when /root/reference is absent (GPU box)."""
    if not os.path.isfile(os.path.join(SRC, "models", "nets", "lvd.py")):
        return DST if os.path.isfile(os.path.join(DST, "models", "nets", "lvd.py")) else None
    manifest = []
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest.append(f"{_sha(d)}  {rel}")
    for rel in DIRS:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(s, d)
    pkg = os.path.join(DST, "tools", "__init__.py")
    if not os.path.exists(pkg) and os.path.exists(os.path.join(SRC, "tools", "__init__.py")):
        shutil.copyfile(os.path.join(SRC, "tools", "__init__.py"), pkg)
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as f:
        f.write("# byte-for-byte copies of files under the reference root, staged by oracle/build_ref.py\n")
        f.write("\n".join(manifest) + "\n")
    if verbose:
        print(f"[build_ref] staged {len(manifest)} files + {DIRS} under {DST}")
    return DST


if __name__ == "__main__":
    p = build(verbose=True)
    if p is None:
        print("[build_ref] no reference available", file=sys.stderr)
        sys.exit(1)
