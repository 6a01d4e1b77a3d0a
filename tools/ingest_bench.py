"""Ingest pipeline check: a .bed-shaped payload on disk / in the page cache -> RheEngine through
`load_genotypes_async` (pinned staging ring + worker threads + side-stream copies) against the serial loader.
Usage (GPU box): python tools/ingest_bench.py [N M J]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyrhe_b200 import synth                      # noqa: E402
from pyrhe_b200.assemble import PathPlan          # noqa: E402
from pyrhe_b200.engine import RheEngine           # noqa: E402
from pyrhe_b200.hostmath import host_terms        # noqa: E402


def main():
    N, M, J = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (200_000, 40_000, 8)
    K, B = 8, 10
    rng = np.random.default_rng(0)
    pitch = (N + 3) // 4
    path = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp", "pyrhe_b200_ingest.npy")
    arr = np.lib.format.open_memmap(path, mode="w+", dtype=np.uint8, shape=(M, pitch))
    blk = rng.integers(0, 256, size=(1000, pitch), dtype=np.uint8) & 0xEE      # no missing code (01) in any pair
    for i in range(0, M, 1000):
        arr[i:i + 1000] = blk[: min(1000, M - i)]
    arr.flush()
    del arr
    try:
        packed = np.load(path, mmap_mode="r")
        annot = synth.random_annot(M, K, rng)
        plan = PathPlan(model="rhe", K=K, B=B, C=0, Ty=1)
        eng = RheEngine(plan, n_indv=N, keep=np.ones(N, bool), annot=annot, num_jack=J, impute="binary", seed=0,
                        device=torch.device("cuda", 0))
        Z = rng.standard_normal((N, B))
        y = rng.standard_normal((N, 1))
        y -= y.mean()
        _, Yres = host_terms(plan, Z, None, y, None)
        eng.set_rhs(Z, None, Yres, None)
        gb = M * pitch / 1e9
        out = None
        for w in (1, 4, 8):
            t0 = time.perf_counter()
            out = eng.run(upload=eng.load_genotypes_async(packed, n_workers=w))
            dt = time.perf_counter() - t0
            print(f"pipelined ingest, {w} staging threads: {dt:.3f} s = {gb / dt:.2f} GB/s of .bed payload ({gb:.1f} GB)")
        t0 = time.perf_counter()
        eng.load_genotypes(packed)
        out2 = eng.run()
        dt = time.perf_counter() - t0
        print(f"serial pageable upload, then compute: {dt:.3f} s = {gb / dt:.2f} GB/s")
        print("same result:", bool(np.allclose(out["XX"], out2["XX"], rtol=1e-6)))
        eng.close()
    finally:
        os.remove(path)


if __name__ == "__main__":
    main()
