"""Put the UNMODIFIED reference package where it can travel to the GPU box: ``baseline/_ref/`` (git-ignored, not
gpurun-ignored).

    python tools/install_reference.py [--source /root/reference]

Outcome of the offline pip route (recorded in DESIGN.md section 5): ``python -m pip install --no-index
--no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference>`` succeeds
but installs ONLY ``cusrl/__init__.py`` and ``cusrl/__main__.py`` -- the reference's ``pyproject.toml`` lists
``packages = ["cusrl"]`` without its sub-packages, so its wheel is not importable.  The reference is pure Python, so the
install is the package tree itself: ``cusrl/`` is copied byte-for-byte (no file is edited; ``INSTALL.json`` records a
SHA-256 over the tree), next to two stand-in modules for its pure-Python dependencies that this image lacks
(``gymnasium``: annotations only on this path, ``objprint``: repr decorator), which are put on ``sys.path`` only when
the real packages are missing (``reference_path()`` below).

Nothing under ``baseline/_ref`` is product code: ``bench.py --impl reference`` / ``reference_cuda`` and the
reference-boundary tests are its only users.
"""

from __future__ import annotations

import argparse
import hashlib
import importlib.util
import json
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
DEST = ROOT / "baseline" / "_ref"
SHIMS_SRC = ROOT / "tests" / "golden" / "_shims"


def tree_sha256(root: Path) -> tuple[str, int]:
    h, n = hashlib.sha256(), 0
    for f in sorted(root.rglob("*.py")):
        h.update(str(f.relative_to(root)).encode())
        h.update(f.read_bytes())
        n += 1
    return h.hexdigest(), n


def install(source: Path = Path("/root/reference"), force: bool = False) -> Path | None:
    """Copy ``<source>/cusrl`` to ``baseline/_ref/cusrl``; returns the install dir, or None when `source` is absent (GPU
    box: the prebuilt copy that travelled with the snapshot is used as is)."""
    pkg = source / "cusrl"
    if not pkg.is_dir():
        return DEST if (DEST / "cusrl").is_dir() else None
    digest, count = tree_sha256(pkg)
    info_file = DEST / "INSTALL.json"
    if not force and info_file.exists() and json.loads(info_file.read_text()).get("sha256") == digest:
        return DEST
    if DEST.exists():
        shutil.rmtree(DEST)
    DEST.mkdir(parents=True)
    shutil.copytree(pkg, DEST / "cusrl", ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    shutil.copytree(SHIMS_SRC, DEST / "_shims", ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    copied, _ = tree_sha256(DEST / "cusrl")
    if copied != digest:
        raise RuntimeError("baseline/_ref/cusrl differs from the reference tree it was copied from")
    info_file.write_text(json.dumps({
        "source": str(source), "sha256": digest, "python_files": count,
        "how": "byte-for-byte copy of the reference's cusrl/ package (pip --target installs only the top-level module: "
               "pyproject.toml omits the sub-packages); _shims/ = stand-ins for gymnasium/objprint, used only if missing",
    }, indent=1))
    return DEST


def reference_path() -> list[str]:
    """sys.path entries that make ``import cusrl`` resolve to the unmodified reference: the travelling copy, else the
    build container's read-only mount; plus the stand-in modules for whichever of gymnasium / objprint is missing."""
    for base, shims in ((DEST, DEST / "_shims"), (Path("/root/reference"), SHIMS_SRC)):
        if (base / "cusrl").is_dir():
            entries = [str(base)]
            if any(importlib.util.find_spec(m) is None for m in ("gymnasium", "objprint")):
                entries.insert(0, str(shims))
            return entries
    raise RuntimeError("the reference package is not available: run `python tools/install_reference.py` in the build "
                       "container (it copies /root/reference/cusrl to baseline/_ref)")


def import_reference():
    """``import cusrl`` (the reference) with the path set up; returns the module."""
    for entry in reversed(reference_path()):
        if entry not in sys.path:
            sys.path.insert(0, entry)
    import cusrl

    return cusrl


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--source", default="/root/reference")
    ap.add_argument("-f", "--force", action="store_true")
    a = ap.parse_args()
    out = install(Path(a.source), a.force)
    print(out if out else "reference source not found and no prebuilt baseline/_ref")
