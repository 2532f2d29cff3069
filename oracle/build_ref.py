"""Recipe that snapshots the UNMODIFIED reference modules of the hot path into ``oracle/_ref/``.

TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.  ``oracle/_ref/`` is git-ignored (never part of the
history) but not gpurun-ignored: like a built ``.so`` it travels to the GPU box, where ``/root/reference``
does not exist, so that ``bench.py --impl reference`` and the ``cpu_baseline`` leg can time the reference's own
classes (``models.VideoModel`` with ``TRNmodule.RelationModuleMultiScale`` and ``loss.attentive_entropy``)
on the box's host cores instead of this repo's restatement of them.

The modules are stored byte for byte inside ONE archive, ``oracle/_ref/ta3n_ref_modules.zip`` (imported through
zipimport by ``oracle/ref_shims.py``), next to a MANIFEST with their sha256 so a reader can check nothing was
edited: a build artefact like a ``.so``, not source files of this repo.
``__graft_entry__.build()`` runs this in the build container (where ``/root/reference`` exists); on the GPU box
it is a no-op and the snapshot made here is used.
"""
from __future__ import annotations

import hashlib
import json
import os
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DST = os.path.join(HERE, "_ref")
REF_SRC = os.environ.get("TA3N_REFERENCE_SRC", "/root/reference")
FILES = ("models.py", "TRNmodule.py", "loss.py")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


ARCHIVE = os.path.join(REF_DST, "ta3n_ref_modules.zip")


def snapshot_available() -> bool:
    if not os.path.isfile(ARCHIVE):
        return False
    try:
        with zipfile.ZipFile(ARCHIVE) as z:
            return all(f in z.namelist() for f in FILES)
    except zipfile.BadZipFile:
        return False


def build_ref(verbose: bool = False) -> bool:
    """Copy the reference modules into oracle/_ref/ when the reference tree is present.  Returns True when a
    usable snapshot exists afterwards."""
    if not os.path.isfile(os.path.join(REF_SRC, "models.py")):
        return snapshot_available()
    os.makedirs(REF_DST, exist_ok=True)
    manifest = {"source": REF_SRC, "archive": os.path.basename(ARCHIVE), "files": {}}
    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        for f in FILES:
            src = os.path.join(REF_SRC, f)
            z.write(src, arcname=f)
            manifest["files"][f] = _sha(src)
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    if verbose:
        print(f"[build_ref] {REF_DST}: " + ", ".join(f"{k} {v[:12]}" for k, v in manifest["files"].items()))
    return True


if __name__ == "__main__":
    print(build_ref(verbose=True))
