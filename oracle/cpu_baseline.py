"""CPU baseline for bench.py: the oracle (a port of the reference algorithm) timed on the host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/rhe_oracle.py).  The reference's own
parallelism is one OS process per contiguous range of jackknife blocks
(/root/reference/pyrhe/src/util/mp_handler.py:27-37, base.py:530-544) with BLAS threads inside
each; this driver mirrors that: `workers` spawned processes, each running the reference block
loop (decode -> impute -> bin gather -> standardise -> XXz / UXXz / XXUz / yXXy, then aggregate)
on its own block of `snps_per_block` SNPs at the FULL number of individuals, with
torch threads = cores // workers.  The full configuration is infeasible on the host
(one 10k-SNP block at N = 500k is 20 GB of float32 plus copies), so the sample is
bounded and throughput is reported per genotype.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np


def _worker(args):
    wid, N, m, K, C, B, threads, repeat = args
    import torch
    torch.set_num_threads(threads)
    from oracle import rhe_oracle
    from pyrhe_b200 import synth
    rng = np.random.default_rng(1000 + wid)
    packed = synth.pack_counts(synth.random_counts(N, m, rng))
    annot = synth.random_annot(m, K, rng)
    Z = rng.standard_normal((N, B))
    W = rng.standard_normal((N, C)) if C else None
    y = rng.standard_normal((N, 1))
    y -= y.mean()
    prob = rhe_oracle.OracleProblem(packed=packed, n_indv_original=N, annot=annot, Z=Z, y=y, num_jack=1, W=W,
                                    impute="binary", seed=0, model="rhe")
    times = []
    for _ in range(repeat):
        o = rhe_oracle.Oracle(prob)
        t0 = time.perf_counter()
        o.pre_compute()
        times.append(time.perf_counter() - t0)
    return times


class CpuBaseline:
    """Persistent pool so that process start-up and `import torch` stay outside the timed region."""

    def __init__(self, N, K, C, B, snps_per_block=200, workers=None):
        cores = os.cpu_count() or 1
        self.cores = cores
        # one block is ~2 GB of fp64 state per worker at N = 500k; cap the pool by memory
        self.workers = workers or max(1, min(cores, 8))
        self.threads = max(1, cores // self.workers)
        self.N, self.K, self.C, self.B, self.m = N, K, C, B, snps_per_block
        self.pool = mp.get_context("spawn").Pool(self.workers)

    def step(self, repeat=1):
        """One bounded sample: every worker processes one block.  Returns (seconds, genotypes)."""
        jobs = [(w, self.N, self.m, self.K, self.C, self.B, self.threads, repeat) for w in range(self.workers)]
        t0 = time.perf_counter()
        res = self.pool.map(_worker, jobs)
        wall = time.perf_counter() - t0
        # data generation inside the worker is not part of the path: use the slowest worker's
        # timed region (workers run concurrently, so this is the wall time of the sample)
        secs = max(sum(r) for r in res)
        return secs, wall, float(self.N) * self.m * self.workers * repeat

    def describe(self):
        return (f"oracle port of the reference block loop, N={self.N}, {self.workers} blocks x {self.m} SNPs "
                f"(one per process), K={self.K}, C={self.C}, B={self.B}, {self.workers} processes x "
                f"{self.threads} torch threads; throughput scales linearly in M")

    def close(self):
        self.pool.close()
        self.pool.join()
