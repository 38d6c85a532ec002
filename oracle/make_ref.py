"""Recipe for `oracle/_ref/`: a verbatim copy of the reference's Python package, made from the sources where they lie.

TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT.  The reference (xqz-u/contrast-gan-3D) is pure Python over ATen, so
"building" it is copying `contrast_gan_3D/**/*.py` (nothing else: no notebooks, no data) from `/root/reference` into the
git-ignored `oracle/_ref/contrast_gan_3D/`.  The copy never enters the history (`.gitignore`: `oracle/_ref/`) but travels
to the GPU box with the snapshot, where `/root/reference` does not exist, so that `bench.py --impl reference` and the
`cpu_baseline` leg can time the UNMODIFIED reference `Trainer.train_step` (imported through `oracle/ref_shim.py`) on the
box's host cores.  Nothing under `contrast_gan_3d_b200/` imports it.

    python -m oracle.make_ref          # (re)create oracle/_ref from /root/reference
"""
from __future__ import annotations

import hashlib
import shutil
import sys
from pathlib import Path

SRC = Path("/root/reference/contrast_gan_3D")
DST = Path(__file__).resolve().parent / "_ref"


def _digest(root: Path) -> str:
    h = hashlib.sha256()
    for f in sorted(root.rglob("*.py")):
        h.update(str(f.relative_to(root)).encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def make(force: bool = False):
    """Copy the reference package; returns the destination, or None when /root/reference is absent (GPU box)."""
    if not SRC.is_dir():
        return DST if (DST / "contrast_gan_3D").is_dir() else None
    dig = _digest(SRC)
    stamp = DST / "SOURCE_SHA256"
    if not force and stamp.exists() and stamp.read_text().strip() == dig:
        return DST
    pkg = DST / "contrast_gan_3D"
    if pkg.exists():
        shutil.rmtree(pkg)
    for f in SRC.rglob("*.py"):
        out = pkg / f.relative_to(SRC)
        out.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(f, out)
    assert _digest(pkg) == dig, "copy differs from the source tree"
    stamp.write_text(dig + "\n")
    return DST


if __name__ == "__main__":
    print(make(force="--force" in sys.argv))
