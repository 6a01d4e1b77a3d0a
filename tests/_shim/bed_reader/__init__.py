"""TEST-ONLY stand-in for the third-party `bed_reader` 1.0.2 wheel (Rust core).

The reference imports `from bed_reader import open_bed`
(/root/reference/pyrhe/src/base/base.py:10) and calls
`open_bed(path).read(index=np.s_[::1, a:b])` (base.py:341,343).  The wheel is
not installed here and there is no network, so `tools/make_golden.py` puts this
directory on PYTHONPATH to run the UNMODIFIED reference.  It follows
bed-reader's documented defaults: dtype float32, order 'F', count_A1=True,
missing -> NaN.  It is never imported by the product.
"""
import numpy as np

# PLINK 2-bit code -> count of A1 (bed-reader default count_A1=True); code 1 = missing
_CODE_TO_A1 = np.array([2.0, np.nan, 1.0, 0.0], dtype=np.float32)


class open_bed:
    def __init__(self, path, *args, **kwargs):
        self.path = str(path)
        with open(self.path, "rb") as f:
            if f.read(3) != bytes([0x6C, 0x1B, 0x01]):
                raise ValueError("not a SNP-major PLINK .bed file")
        stem = self.path[:-4]
        with open(stem + ".fam") as f:
            self.iid_count = sum(1 for _ in f)
        with open(stem + ".bim") as f:
            self.sid_count = sum(1 for _ in f)
        self._row_bytes = (self.iid_count + 3) // 4

    @property
    def shape(self):
        return (self.iid_count, self.sid_count)

    def read(self, index=None, dtype="float32", order="F", **kwargs):
        rows, cols = slice(None), slice(None)
        if index is not None:
            if isinstance(index, tuple):
                rows, cols = index
            else:
                cols = index
        sids = np.arange(self.sid_count)[cols]
        out = np.empty((self.iid_count, len(sids)), dtype=dtype, order=order)
        raw = np.memmap(self.path, dtype=np.uint8, mode="r", offset=3,
                        shape=(self.sid_count, self._row_bytes))
        for c, s in enumerate(sids):
            b = np.asarray(raw[s])
            codes = np.stack([b & 3, (b >> 2) & 3, (b >> 4) & 3, (b >> 6) & 3], axis=1).reshape(-1)
            out[:, c] = _CODE_TO_A1[codes[: self.iid_count]]
        return out[rows]
