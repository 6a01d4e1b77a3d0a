"""Per-phase times of the block path on a few config-5 blocks, with and without the individual-major fast layout,
and a bit-for-bit comparison of the block partials the two pass-B kernels produce.

    python tools/passb_probe.py [workload] [steps]
"""
import ctypes as C
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from pyrhe_b200 import _lib

wl_name = sys.argv[1] if len(sys.argv) > 1 else "profile5"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = dict(bench.WORKLOADS[wl_name])
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
args = types.SimpleNamespace(kernel_path=1, retile=False)
pb = bench.build_problem(wl, args, 0, 1, dev)
eng = pb["eng"]
lib = _lib.load()
print(f"{wl_name}: blocks {len(eng.own)}, fast-layout copies {len(eng.gt)}, HBM genotypes {eng.genotype_bytes() / 1e9:.2f} GB")
names = ["params", "pass_a", "standardize_gram", "pass_b"]
out = {}
for fast in (True, False):
    eng.use_fast_layout = fast
    eng.run()
    _lib.check(lib.rhe_timing_enable(eng._ctx, 1))
    for _ in range(steps):
        eng.run()
    ph = (C.c_double * 4)()
    n = C.c_int32()
    _lib.check(lib.rhe_timing_collect(eng._ctx, ph, C.byref(n)))
    _lib.check(lib.rhe_timing_enable(eng._ctx, 0))
    print("fast" if fast else "gather", {k: round(ph[i] / max(n.value, 1), 4) for i, k in enumerate(names)}, flush=True)
    out[fast] = (eng.P_all.clone() if eng.P_all is not None else None, eng.S.clone())
if out[True][0] is not None:
    a, b = out[True][0], out[False][0]
    print("P identical:", bool(torch.equal(a, b)), "max |diff|:", float((a - b).abs().max()), "max |P|:", float(b.abs().max()))
    nz = torch.nonzero(a != b)
    print("differing entries:", nz.shape[0], "of", a.numel())
    for r in nz[:8].tolist():
        print("  [block, e, b, i] =", r, "fast", float(a[tuple(r)]), "gather", float(b[tuple(r)]))
a, b = out[True][1], out[False][1]
print("S max |diff| / max |S|:", float((a - b).abs().max() / b.abs().max()))

# the same blocks with their rows re-tiled for pass A (contiguous boxes of imputed counts); results must not change
G0 = eng.run()["G_blk"]
eng.close()
del pb, eng
torch.cuda.empty_cache()
pb = bench.build_problem(wl, types.SimpleNamespace(kernel_path=1, retile=True), 0, 1, dev)
eng = pb["eng"]
eng.run()
_lib.check(lib.rhe_timing_enable(eng._ctx, 1))
for _ in range(steps):
    pieces = eng.run()
ph = (C.c_double * 4)()
n = C.c_int32()
_lib.check(lib.rhe_timing_collect(eng._ctx, ph, C.byref(n)))
print("retiled", {k: round(ph[i] / max(n.value, 1), 4) for i, k in enumerate(names)}, "tiled blocks", len(eng._tiled), flush=True)
print("G_blk max |diff| / max:", float(np.abs(pieces["G_blk"] - G0).max() / np.abs(G0).max()),
      "P identical to the PLINK-row run:", bool(torch.equal(eng.P_all, out[True][0])) if out[True][0] is not None else None)
