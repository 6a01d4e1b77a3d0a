"""Synthetic PLINK-1 data sets for tests and benchmarks.

Writes the on-disk formats the reference consumes (`.bed/.bim/.fam`, `.pheno`,
`.cov`, `.env`, annotation) so the same files can be fed to the reference
(through the `bed_reader` test shim) and to this package.

`.bed` layout: magic ``6C 1B 01`` then one row per SNP of
``ceil(N/4)`` bytes, four genotypes per byte LSB-first, codes
``00`` hom-A1, ``01`` missing, ``10`` het, ``11`` hom-A2.  After the flip at
/root/reference/pyrhe/src/base/base.py:352-355 the value the path works with is
the A2 count: ``00->0, 10->1, 11->2, 01->missing``.
"""
from __future__ import annotations

import os
import numpy as np

BED_MAGIC = bytes([0x6C, 0x1B, 0x01])
# A2 allele count (0,1,2) or missing (3)  ->  2-bit PLINK code
_COUNT_TO_CODE = np.array([0b00, 0b10, 0b11, 0b01], dtype=np.uint8)


def pack_counts(counts: np.ndarray) -> np.ndarray:
    """counts: [M, N] uint8 in {0,1,2,3(missing)} -> packed rows [M, ceil(N/4)] uint8."""
    M, N = counts.shape
    nb = (N + 3) // 4
    codes = _COUNT_TO_CODE[counts]
    pad = nb * 4 - N
    if pad:
        codes = np.concatenate([codes, np.zeros((M, pad), np.uint8)], axis=1)
    codes = codes.reshape(M, nb, 4)
    return (codes[:, :, 0] | (codes[:, :, 1] << 2) | (codes[:, :, 2] << 4) | (codes[:, :, 3] << 6)).astype(np.uint8)


def random_counts(N: int, M: int, rng: np.random.Generator, missing_rate: float = 0.0,
                  maf_lo: float = 0.05, maf_hi: float = 0.5) -> np.ndarray:
    """[M, N] uint8 A2 counts ~ Binomial(2, p_s), p_s ~ U(maf_lo, maf_hi); 3 marks missing."""
    p = rng.uniform(maf_lo, maf_hi, size=(M, 1))
    g = (rng.random((M, N)) < p).astype(np.uint8) + (rng.random((M, N)) < p).astype(np.uint8)
    if missing_rate > 0:
        g[rng.random((M, N)) < missing_rate] = 3
    return g


def write_bed(path: str, counts: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(BED_MAGIC)
        f.write(pack_counts(counts).tobytes())


def write_plink_text(prefix: str, N: int, M: int) -> None:
    with open(prefix + ".fam", "w") as f:
        for i in range(N):
            f.write(f"{i} {i} 0 0 0 -9\n")
    with open(prefix + ".bim", "w") as f:
        for s in range(M):
            f.write(f"1\trs{s}\t0\t{s}\tA\tG\n")


def write_pheno(path: str, Y: np.ndarray, missing_rows=()) -> None:
    N, T = Y.shape
    miss = set(int(i) for i in missing_rows)
    with open(path, "w") as f:
        f.write("FID IID " + " ".join(f"pheno{t}" for t in range(T)) + "\n")
        for i in range(N):
            vals = ["NA"] * T if i in miss else [repr(float(v)) for v in Y[i]]
            f.write(f"{i} {i} " + " ".join(vals) + "\n")


def write_cov(path: str, W: np.ndarray, missing_cells=()) -> None:
    N, C = W.shape
    miss = set((int(i), int(c)) for i, c in missing_cells)
    with open(path, "w") as f:
        f.write("FID IID " + " ".join(f"cov{c}" for c in range(C)) + "\n")
        for i in range(N):
            f.write(f"{i} {i} " + " ".join("NA" if (i, c) in miss else repr(float(W[i, c])) for c in range(C)) + "\n")


def write_env(path: str, env: np.ndarray) -> None:
    with open(path, "w") as f:
        f.write("FID IID env\n")
        for i, v in enumerate(env):
            f.write(f"{i} {i} {int(v) if float(v).is_integer() else repr(float(v))}\n")


def write_annot(path: str, annot: np.ndarray) -> None:
    with open(path, "w") as f:
        for row in annot:
            f.write(" ".join(str(int(v)) for v in row) + "\n")


def random_annot(M: int, K: int, rng: np.random.Generator, overlap: float = 0.0) -> np.ndarray:
    """One bin per SNP (as file_processing.py:109-118); `overlap` adds a second bin to that fraction."""
    annot = np.zeros((M, K), dtype=np.int64)
    annot[np.arange(M), rng.integers(0, K, size=M)] = 1
    if overlap > 0 and K > 1:
        extra = np.nonzero(rng.random(M) < overlap)[0]
        annot[extra, rng.integers(0, K, size=extra.size)] = 1
    return annot


def make_dataset(outdir: str, name: str, N: int, M: int, K: int, *, seed: int = 0,
                 n_cov: int = 0, n_traits: int = 1, missing_rate: float = 0.0,
                 missing_pheno=(), with_env: bool = False, h2: float = 0.25,
                 overlap: float = 0.0, maf_lo: float = 0.05, maf_hi: float = 0.5,
                 monomorphic=(), binary_pheno: bool = False, cov_missing=()) -> dict:
    """Write a full synthetic data set; returns the paths (keys = reference kwarg names).

    The later keywords only change the data when given (the random stream of the existing cases is untouched):
    `maf_lo/maf_hi` the allele-frequency range (rare variants), `monomorphic` SNP rows forced to count 0,
    `binary_pheno` thresholds every trait at its median into {0, 1} (case/control, rhe.py:79-88), `cov_missing`
    (row, column) covariate cells written as NA."""
    os.makedirs(outdir, exist_ok=True)
    rng = np.random.default_rng(seed)
    prefix = os.path.join(outdir, name)
    counts = random_counts(N, M, rng, missing_rate, maf_lo, maf_hi)
    for s in monomorphic:
        counts[s] = 0
    write_bed(prefix + ".bed", counts)
    write_plink_text(prefix, N, M)
    annot = random_annot(M, K, rng, overlap)
    write_annot(prefix + ".annot", annot)
    # phenotype with a genetic component so the estimates are not degenerate
    g = np.where(counts == 3, 0, counts).astype(np.float64).T  # [N, M]
    g = (g - g.mean(0)) / np.maximum(g.std(0), 1e-9)
    Y = np.empty((N, n_traits))
    for t in range(n_traits):
        beta = rng.standard_normal(M) * np.sqrt(h2 / M)
        Y[:, t] = g @ beta + rng.standard_normal(N) * np.sqrt(1 - h2)
    paths = {"geno_file": prefix, "annot_file": prefix + ".annot", "pheno_file": prefix + ".pheno"}
    if n_cov:
        W = rng.standard_normal((N, n_cov))
        W[:, 0] = (rng.random(N) < 0.5).astype(np.float64)
        Y += (W @ rng.standard_normal((n_cov, 1))) * 0.3
        write_cov(prefix + ".cov", W, cov_missing)
        paths["cov_file"] = prefix + ".cov"
    if with_env:
        env = (rng.random(N) < 0.4).astype(np.float64)
        write_env(prefix + ".env", env)
        paths["env_file"] = prefix + ".env"
    if binary_pheno:
        Y = (Y > np.median(Y, axis=0)).astype(np.float64)
    write_pheno(prefix + ".pheno", Y, missing_pheno)
    return paths
