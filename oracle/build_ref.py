"""ORACLE SUPPORT (test / baseline infrastructure, NOT product code): recipe that makes the REFERENCE'S OWN hot-path files
available next to the oracle as `oracle/_ref/` (git-ignored -- reference sources never enter the history -- but not
gpurun-ignored, so the copy travels to the GPU box where /root/reference does not exist).

    python -m oracle.build_ref            # copies from /root/reference when it is mounted; no-op otherwise

The reference is pure Python (no native code, no build), so "building" it is copying the files its hot path
imports, unmodified, and recording their sha256 in oracle/_ref/MANIFEST.json. Used by:
  * bench.py --impl reference              : the reference's own modules stepping on the host CPU cores (kind "reference")
  * bench.py --impl reference-gpu-eager    : the same unmodified modules on one B200 under fp16 autocast (trainer.py:126)
  * tests/test_reference_trainer_gpu.py    : the reference's unmodified `missing_trainer` driving the B200 model
  * tools/make_golden.py uses /root/reference directly (build container only).
Import shims live in oracle/ref_loader.py; nothing in _ref is edited.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("TMP_REFERENCE_ROOT", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")

# every file the path `get_model(args)(args)` + `builder.trainer.get_trainer` imports
FILES = [
    "control/config.py",
    "builder/models/__init__.py",
    "builder/models/8_missing_models/tri_mbt_vsltcls.py",
    "builder/models/src/__init__.py",
    "builder/models/src/swin_transformer.py",
    "builder/models/src/vision_transformer.py",
    "builder/models/src/reports_transformer_decoder.py",
    "builder/models/src/transformer/__init__.py",
    "builder/models/src/transformer/attention.py",
    "builder/models/src/transformer/encoder.py",
    "builder/models/src/transformer/mbt_encoder.py",
    "builder/models/src/transformer/module.py",
    "builder/models/src/transformer/utils.py",
    "builder/trainer/__init__.py",
    "builder/trainer/trainer.py",
    "builder/utils/__init__.py",
    "builder/utils/cosine_annealing_with_warmup_v2.py",
]


def build(verbose: bool = False) -> str | None:
    """Copy the reference hot-path files into oracle/_ref/. Returns the path, or None when the reference is not mounted
    (the GPU box: the copy made in the build container is used)."""
    if not os.path.isdir(REF_SRC):
        return REF_DST if os.path.exists(os.path.join(REF_DST, "MANIFEST.json")) else None
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF_SRC, rel)
        dst = os.path.join(REF_DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(src, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_SRC, "files": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(FILES)} reference files copied from {REF_SRC}")
    return REF_DST


def available() -> bool:
    return os.path.exists(os.path.join(REF_DST, "MANIFEST.json"))


def verify() -> bool:
    """sha256 of every copied file against the manifest (the copy is unmodified)."""
    with open(os.path.join(REF_DST, "MANIFEST.json")) as f:
        man = json.load(f)["files"]
    for rel, h in man.items():
        with open(os.path.join(REF_DST, rel), "rb") as f:
            if hashlib.sha256(f.read()).hexdigest() != h:
                return False
    return True


if __name__ == "__main__":
    p = build(verbose=True)
    print(p or "reference not mounted and no oracle/_ref present")
    sys.exit(0)
