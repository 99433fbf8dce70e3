"""
ORACLE support (test infrastructure): place an UNMODIFIED copy of the reference files the hot path touches under
oracle/_ref/ so that they can travel to the GPU box, where /root/reference does not exist.

    python -m oracle.make_ref          # no-op when /root/reference is absent

oracle/_ref/ is git-ignored (the reference's sources never enter this repository's history) but not
gpurun-ignored.  Consumers:
  * bench.py --impl reference     times the reference's own GibbsSampler.gibbs_sweep on the box's host cores
  * tests/test_reference_suite.py runs the reference's own tests/test_{gibbs,ising,core}.py and benchmark drivers
                                  against the tsu import shim (tsu_emulator_b200/compat)
Files are byte-for-byte copies; oracle/_ref/MANIFEST.json records their sha256.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("TSU_REFERENCE_SOURCE", "/root/reference")

FILES = [
    "tsu/gibbs.py",
    "tsu/core.py",
    "tsu/models/__init__.py",
    "tsu/models/ising.py",
    "tsu/api.py",
    "tsu/benchmarks/__init__.py",
    "tsu/benchmarks/sampling.py",
    "tsu/benchmarks/comparison.py",
    "tsu/benchmarks/optimization.py",
    "tests/test_gibbs.py",
    "tests/test_ising.py",
    "tests/test_core.py",
]


def make_ref(verbose: bool = False) -> bool:
    if not os.path.isfile(os.path.join(SOURCE, "tsu", "gibbs.py")):
        return os.path.isfile(os.path.join(DEST, "tsu", "gibbs.py"))
    manifest = {}
    for rel in FILES:
        src = os.path.join(SOURCE, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
        if verbose:
            print("copied", rel)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SOURCE, "sha256": manifest}, fh, indent=1)
    return True


def ref_root():
    """directory that holds the reference tree: the live one in the build container, the copy on the GPU box"""
    if os.path.isfile(os.path.join(SOURCE, "tsu", "gibbs.py")):
        return SOURCE
    if os.path.isfile(os.path.join(DEST, "tsu", "gibbs.py")):
        return DEST
    return None


if __name__ == "__main__":
    ok = make_ref(verbose=True)
    print("oracle/_ref ready" if ok else "reference tree not found; nothing copied")
    sys.exit(0)
