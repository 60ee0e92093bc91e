"""Recipe for ``oracle/_ref``: the reference's own hot-path files, staged for the GPU box.  TEST INFRASTRUCTURE.

``/root/reference`` exists only in the authoring container.  The reference is pure Python, so "building" it means
staging the few files of the hot path under ``oracle/_ref/`` -- a build output, exactly like
``liblpb200.so``: listed in ``.gitignore`` (reference sources never enter the history), not in ``.gpurunignore`` (it
travels to the GPU box).  ``oracle/ref_loader.py`` executes them unmodified from there when ``/root/reference`` is
absent, which is what lets ``bench.py --impl reference`` and ``cpu_baseline`` time the reference's own code
(``kind: "reference"``) next to the GPU numbers.  ``__graft_entry__.build()`` runs this when the reference is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

SRC_ROOT = os.environ.get("LP_REFERENCE_ROOT", "/root/reference")
DST_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

# hot-path files only; everything they import beyond these is stubbed by ref_loader
FILES = [
    "lightning_pose/models/heads/heatmap.py",
    "lightning_pose/models/heads/heatmap_mhcrnn.py",
    "lightning_pose/models/backbones/factory.py",
    "lightning_pose/data/heatmaps.py",
    "lightning_pose/data/utils.py",
    "lightning_pose/data/bboxes.py",
    "lightning_pose/data/datatypes.py",  # TypedDicts imported by bboxes.py
    "lightning_pose/losses/losses.py",
    "lightning_pose/losses/factory.py",
    "lightning_pose/utils/pca.py",
]


def build(verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(SRC_ROOT, "lightning_pose")):
        if verbose:
            print(f"oracle/build_ref: {SRC_ROOT} absent - keeping the staged copy" if os.path.isdir(DST_ROOT) else "oracle/build_ref: no reference here")
        return os.path.isdir(DST_ROOT)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC_ROOT, rel), os.path.join(DST_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DST_ROOT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC_ROOT, "files": manifest}, fh, indent=1)
    if verbose:
        print(f"oracle/build_ref: staged {len(FILES)} reference files under {DST_ROOT}")
    return True


if __name__ == "__main__":
    build()
