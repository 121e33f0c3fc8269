#!/usr/bin/env python
"""Stage the UNMODIFIED reference for the baseline arms of bench.py (`--impl reference`, `--impl reference-cuda`).

The reference is pure Python without a setup.py / pyproject (pip cannot install it), so "installing" it is copying the
files of the hot path, byte for byte, from /root/reference into baseline/_ref/ (git-ignored -- reference sources never enter
this repository's history -- but not gpurun-ignored, so the copy travels to the GPU box, where /root/reference does not
exist).  Copied: wild_completion/*.py, deepsdf/deep_sdf/*.py, deepsdf/networks/*.py, metrics_3d/*.py, the host scripts (*.py at the top
level), configs/*.yaml and, of the two
shipped models, specs.json + ModelParameters/latest.pth + LatentCodes/latest.pth (OptimizerParameters is training state the
reference never reads).  A MANIFEST with sha256 sums is written next to them so a run can state what it timed.

    python scripts/vendor_reference.py [/root/reference]
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
PATTERNS = [("wild_completion", ".py"), ("deepsdf/deep_sdf", ".py"), ("deepsdf/networks", ".py"), ("deepsdf", ".py"),
            ("metrics_3d", ".py"), ("configs", ".yaml"), (".", ".py")]      # "." = the host scripts (test_wild_completion.py, ...)
MODEL_FILES = ["specs.json", "ModelParameters/latest.pth", "LatentCodes/latest.pth"]


def vendor(src: str = "/root/reference") -> bool:
    if not os.path.isdir(src):
        return False
    manifest = {}

    def cp(rel):
        a, b = os.path.join(src, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(b), exist_ok=True)
        shutil.copyfile(a, b)
        manifest[rel] = hashlib.sha256(open(b, "rb").read()).hexdigest()

    for d, ext in PATTERNS:
        full = os.path.join(src, d)
        if not os.path.isdir(full):
            continue
        for name in sorted(os.listdir(full)):
            if name.endswith(ext) and os.path.isfile(os.path.join(full, name)):
                cp(os.path.normpath(os.path.join(d, name)))
    for model in ("sweetpepper_32", "strawberry_32"):
        for f in MODEL_FILES:
            rel = os.path.join("deepsdf", "models", model, f)
            if os.path.isfile(os.path.join(src, rel)):
                cp(rel)
    for extra in ("README.md", "LICENSE"):
        if os.path.isfile(os.path.join(src, extra)):
            cp(extra)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return True


if __name__ == "__main__":
    ok = vendor(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("vendored into", DST if ok else "(nothing: source missing)")
