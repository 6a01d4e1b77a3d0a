"""Recipe for `oracle/_ref/`: the UNMODIFIED reference as the CPU arm of bench.py.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  sriramlab/PyRHE is pure Python, so there is nothing to compile: the recipe
copies the reference's own package from where it lies (`/root/reference/pyrhe`) into `oracle/_ref/pyrhe` -- outputs only,
git-ignored, shipped to the GPU box with the snapshot (`/root/reference` does not exist there) -- together with the
`bed_reader` stand-in of tests/_shim (the reference imports `bed_reader==1.0.2`, a Rust wheel that is not installed in
this image; the stand-in follows its documented defaults) and a manifest of SHA-256 digests so that
`oracle/ref_baseline.py` can check that what it runs is byte for byte what the reference ships.

    python oracle/make_ref.py          (also run by __graft_entry__.build() when /root/reference is present)
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
OUT = os.path.join(ROOT, "oracle", "_ref")


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def build(verbose=False):
    src = os.path.join(REFERENCE, "pyrhe")
    if not os.path.isdir(src):
        return None                                      # GPU box: only the prebuilt copy is used
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    shutil.copytree(src, os.path.join(OUT, "pyrhe"), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    shutil.copytree(os.path.join(ROOT, "tests", "_shim", "bed_reader"), os.path.join(OUT, "bed_reader"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    manifest = {}
    for base, _, files in os.walk(os.path.join(OUT, "pyrhe")):
        for f in files:
            p = os.path.join(base, f)
            rel = os.path.relpath(p, OUT)
            manifest[rel] = sha256(p)
            assert manifest[rel] == sha256(os.path.join(REFERENCE, rel)), rel
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest, "bed_reader": "tests/_shim/bed_reader (stand-in for bed_reader==1.0.2)"},
                  f, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} reference files copied unmodified")
    return OUT


if __name__ == "__main__":
    if build(verbose=True) is None:
        sys.exit("no /root/reference here")
