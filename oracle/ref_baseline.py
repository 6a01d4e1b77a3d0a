"""CPU arm of bench.py with `kind = "reference"`: the UNMODIFIED reference (oracle/_ref, see oracle/make_ref.py) timed on
the host cores on a bounded sample of the benchmark's workload.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The sample is a real PLINK data set at the FULL number of individuals with
`blocks x snps_per_block` SNPs (synthetic genotypes, 8 bins, covariates, binary imputation -- the workload's shape with
fewer SNPs; throughput is per genotype and the reference's cost is linear in M), written once to local storage; a step
runs `RHE(...)(trait=0)` of the reference in a fresh child process with the reference's own parallelism -- one worker
process per contiguous range of jackknife blocks (`multiprocessing=True`, mp_handler.py:27-37, base.py:530-544) and the
remaining cores as torch / BLAS threads inside each worker -- and takes the seconds of its `pre_compute` (the per-block
path: read_geno, impute, bins, standardise, XXz / UXXz / XXUz / yXXy, aggregate) from the child.  The rest of a user call
(`run()`: the J + 1 normal equations in O(J E^2 B N) host loops, solves, h2) takes three times as long again on such a
small sample and does not grow with M, so the default step leaves it out (`full_call=True` runs it).  `.bed` decoding goes through the numpy stand-in for `bed_reader` (a few percent of the call).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> str:
    """'' when oracle/_ref holds an intact copy of the reference, else the reason."""
    man = os.path.join(REF, "MANIFEST.json")
    if not os.path.exists(man):
        return "oracle/_ref is missing (python oracle/make_ref.py needs /root/reference)"
    files = json.load(open(man))["files"]
    for rel, digest in files.items():
        p = os.path.join(REF, rel)
        if not os.path.exists(p) or hashlib.sha256(open(p, "rb").read()).hexdigest() != digest:
            return f"oracle/_ref/{rel} differs from the reference"
    return ""


class RefBaseline:
    def __init__(self, N, K, C, B, snps_per_block=100, blocks=8, workers=None, full_call=False):
        sys.path.insert(0, os.path.dirname(HERE))
        from pyrhe_b200 import synth
        self.N, self.K, self.C, self.B, self.m, self.blocks = N, K, C, B, snps_per_block, blocks
        self.cores = os.cpu_count() or 1
        self.workers = max(1, min(workers or self.cores, blocks))
        self.threads = max(1, self.cores // self.workers)
        M = blocks * snps_per_block
        root = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 4e9 else None
        self.dir = tempfile.mkdtemp(prefix="pyrhe_ref_", dir=root)
        rng = np.random.default_rng(7)
        prefix = os.path.join(self.dir, "sample")
        with open(prefix + ".bed", "wb") as f:             # block by block: the counts of one block are 100 MB
            f.write(synth.BED_MAGIC)
            for _ in range(blocks):
                f.write(synth.pack_counts(synth.random_counts(N, snps_per_block, rng, missing_rate=0.001)).tobytes())
        ids = np.arange(N)
        np.savetxt(prefix + ".fam", np.stack([ids, ids, 0 * ids, 0 * ids, 0 * ids, 0 * ids - 9], 1), fmt="%d")
        with open(prefix + ".bim", "w") as f:
            f.write("".join(f"1\trs{i}\t0\t{i}\tA\tG\n" for i in range(M)))
        synth.write_annot(prefix + ".annot", synth.random_annot(M, K, rng))
        y = rng.standard_normal(N)
        with open(prefix + ".pheno", "w") as f:
            f.write("FID IID pheno0\n")
            f.write("".join(f"{i} {i} {v!r}\n" for i, v in enumerate(y.tolist())))
        self.paths = dict(geno_file=prefix, annot_file=prefix + ".annot", pheno_file=prefix + ".pheno")
        if C:
            W = rng.standard_normal((N, C))
            W[:, 0] = rng.random(N) < 0.5
            with open(prefix + ".cov", "w") as f:
                f.write("FID IID " + " ".join(f"cov{c}" for c in range(C)) + "\n")
                np.savetxt(f, np.concatenate([ids[:, None], ids[:, None], W], 1), fmt=["%d", "%d"] + ["%.17g"] * C)
            self.paths["cov_file"] = prefix + ".cov"
        self.spec = os.path.join(self.dir, "spec.json")
        json.dump(dict(paths=self.paths, threads=self.threads, workers=self.workers, full_call=bool(full_call),
                       kwargs=dict(num_jack=blocks, num_random_vec=B, seed=0, geno_impute_method="binary")),
                  open(self.spec, "w"))
        self.last = None

    def step(self):
        """One run of the reference on the sample.  Returns (seconds of its pre_compute -- the per-block path, what the
        oracle-port baseline times as well --, seconds of the whole model(trait=0) call, genotypes)."""
        env = dict(os.environ, PYTHONPATH=REF, OMP_NUM_THREADS=str(self.threads), MKL_NUM_THREADS=str(self.threads))
        env.pop("CUDA_VISIBLE_DEVICES", None)
        res = subprocess.run([sys.executable, os.path.join(HERE, "_ref_run.py"), self.spec], cwd=REF, env=env,
                             capture_output=True, text=True, timeout=3600)
        if res.returncode != 0:
            raise RuntimeError("the reference failed on the sample:\n" + res.stderr[-2000:])
        self.last = json.loads(res.stdout.strip().splitlines()[-1])
        return self.last["pre_compute_s"], self.last["call_s"], float(self.N) * self.m * self.blocks

    def describe(self):
        return (f"UNMODIFIED reference (oracle/_ref): RHE(...).pre_compute(), N={self.N}, {self.blocks} jackknife blocks x "
                f"{self.m} SNPs, K={self.K}, C={self.C}, B={self.B}, binary imputation, {self.workers} worker processes "
                f"(the reference's multiprocessing=True) x {self.threads} torch threads, bed_reader stand-in; seconds of "
                f"its pre_compute (the per-block path); throughput per genotype, linear in M")

    def close(self):
        shutil.rmtree(self.dir, ignore_errors=True)
