"""The CPU arm of bench.py: `oracle/_ref` is the reference byte for byte (manifest of SHA-256 digests written by
oracle/make_ref.py) and the unmodified reference runs on a tiny sample through oracle/ref_baseline.py -- with its own
worker processes -- in this container.  Skipped where oracle/_ref has not been built (no /root/reference)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _ref():
    from oracle import make_ref, ref_baseline
    if not os.path.isdir(make_ref.OUT) and os.path.isdir(make_ref.REFERENCE):
        make_ref.build()
    why = ref_baseline.available()
    if why:
        pytest.skip(why)
    return ref_baseline


def test_shipped_copy_is_the_reference_byte_for_byte():
    import hashlib
    import json
    rb = _ref()
    man = json.load(open(os.path.join(rb.REF, "MANIFEST.json")))
    assert len(man["files"]) >= 20 and "pyrhe/src/base/base.py" in man["files"]
    if os.path.isdir("/root/reference"):
        for rel, digest in man["files"].items():
            assert hashlib.sha256(open(os.path.join("/root/reference", rel), "rb").read()).hexdigest() == digest, rel


@pytest.mark.parametrize("workers", [1, 2])
def test_unmodified_reference_runs_a_sample(workers):
    rb = _ref()
    b = rb.RefBaseline(400, 2, 1, 3, snps_per_block=30, blocks=2, workers=workers)
    try:
        secs, _, genotypes = b.step()
        assert secs > 0 and genotypes == 400 * 60
        assert b.last["num_indv"] == 400 and b.last["num_snp"] == 60
        assert "UNMODIFIED reference" in b.describe()
    finally:
        b.close()
