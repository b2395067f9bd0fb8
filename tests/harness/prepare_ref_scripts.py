"""Recipe: place the reference's two training scripts, byte for byte, where the GPU box can run them.

    python tests/harness/prepare_ref_scripts.py                (also run by __graft_entry__.build())

``/root/reference`` exists only in the build container.  The scripts are copied UNCHANGED into the git-ignored
``tests/_ref_scripts/`` (never committed -- like ``oracle/_ref/`` it travels to the GPU box with the gpurun snapshot)
together with a manifest of their SHA-256 digests, so that tests/test_ref_scripts_gpu.py can show the file it ran is
the reference's own (SURVEY.md 8b "Script harness"; north_star: the scripts "run unchanged against it").
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(os.path.dirname(HERE), "_ref_scripts")
REFERENCE = os.environ.get("B200SURV_REFERENCE", "/root/reference")
SCRIPTS = ("scripts/training/simple_fusion.py", "scripts/training/partial_modality_training.py")


def prepare() -> str | None:
    if not os.path.isdir(REFERENCE):
        return DEST if os.path.exists(os.path.join(DEST, "MANIFEST.json")) else None
    os.makedirs(DEST, exist_ok=True)
    manifest = {}
    for rel in SCRIPTS:
        src = os.path.join(REFERENCE, rel)
        dst = os.path.join(DEST, os.path.basename(rel))
        shutil.copyfile(src, dst)
        with open(src, "rb") as fh:
            manifest[os.path.basename(rel)] = {"source": rel, "sha256": hashlib.sha256(fh.read()).hexdigest()}
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    return DEST


if __name__ == "__main__":
    print(prepare())
