"""Host-side estimator tail in fp64 numpy: solve, jackknife SE, h2, enrichment, trace files.

Semantics follow /root/reference/pyrhe/src/base/base.py:298-312,680-868 and
models/genie/genie.py:146-219 (tiny (E+1)^2 problems -- SURVEY.md §2 row 5 keeps them on the
host so that parity is decided by the hot path alone).  Loops over SNPs in the reference's
`compute_h2_overlapping` are replaced by bin co-occurrence counts.
"""
from __future__ import annotations

import os

import numpy as np
import scipy.linalg


def solve_lstsq(T, q):
    """base.py:298-303."""
    return np.linalg.lstsq(T, q, rcond=None)[0]


def solve_qr(T, q):
    """base.py:306-312."""
    Qm, R = scipy.linalg.qr(T)
    return scipy.linalg.solve_triangular(R, np.dot(Qm.T, q))


def solve(T, q, method):
    if method == "lstsq":
        return np.ravel(solve_lstsq(T, q))
    if method == "QR":
        return np.ravel(solve_qr(T, q))
    raise ValueError("Unsupported method for solving linear equation")


def jackknife_se(ests: np.ndarray, num_jack: int) -> list:
    """sqrt((J-1)/J * sum_j (theta_j - mean)^2) per column -- base.py:680-703."""
    ests = np.asarray(ests, dtype=np.float64)
    dev = ests - ests.mean(axis=0)
    return list(np.sqrt((num_jack - 1) * np.sum(dev * dev, axis=0) / num_jack))


def h2_nonoverlapping(sigma_all: np.ndarray, E: int) -> np.ndarray:
    """Rows = jackknife samples (+ total): [h2_1 .. h2_E, h2_SNP] -- base.py:705-742."""
    sigma_all = np.asarray(sigma_all, dtype=np.float64)
    gen = sigma_all[:, :E].sum(axis=1)
    denom = gen + sigma_all[:, -1]
    return np.concatenate([sigma_all[:, :-1] / denom[:, None], (gen / denom)[:, None]], axis=1)


def bin_cooccurrence(annot: np.ndarray, ranges) -> tuple:
    """(total, per-sample-excluded) K x K counts of SNPs carrying both bin a and bin b (value == 1).

    Entry j < J of the second array covers block j.  Entry J reproduces a reference quirk:
    `_get_annot_subsample(J)` (base.py:382-393) masks rows [J * (M // J), M), so the "all SNPs"
    sample of `compute_h2_overlapping` silently drops the M % J remainder SNPs."""
    A = (annot == 1).astype(np.float64)
    J = len(ranges)
    spans = list(ranges) + [(J * (annot.shape[0] // J), annot.shape[0])]
    excluded = np.array([A[a:b].T @ A[a:b] for a, b in spans])
    return A.T @ A, excluded


def h2_overlapping(sigma_all, M_table, cooc_total, cooc_block, E: int) -> np.ndarray:
    """base.py:744-785: a SNP's variance is the sum of sigma_b / M_b over the bins b it
    belongs to; bin k's h2 sums that over its SNPs.  With C[k,b] = #SNPs in both k and b
    (leave-one-block-out) this is C @ (sigma / M)."""
    sigma_all = np.asarray(sigma_all, dtype=np.float64)
    J = sigma_all.shape[0] - 1
    out = []
    for j in range(J + 1):
        s = sigma_all[j]
        Mj = np.asarray(M_table[j], dtype=np.float64)
        per_snp = np.divide(s[:E], Mj[:E], out=np.zeros(E), where=Mj[:E] != 0)
        C = cooc_total - cooc_block[j]
        gen = s[:E].sum()
        denom = gen + s[-1]
        out.append(np.concatenate([(C @ per_snp) / denom, [gen / denom]]))
    return np.array(out)


def enrichment(h2_all: np.ndarray, M_table, E: int) -> np.ndarray:
    """(h2_k / h2_SNP) / (M_k / M) per jackknife sample -- base.py:788-825."""
    h2_all = np.asarray(h2_all, dtype=np.float64)
    out = np.zeros((h2_all.shape[0], E))
    for j in range(h2_all.shape[0]):
        Mj = np.asarray(M_table[j], dtype=np.float64)
        tot = Mj.sum()
        for k in range(E):
            if tot != 0 and Mj[k] != 0:
                out[j, k] = (h2_all[j, k] / h2_all[j, -1]) / (Mj[k] / tot)
    return out


def liability_h2(h2, se, samp_prev, pop_prev):
    """Observed -> liability scale (base.py:857-868)."""
    from scipy.stats import chi2, norm
    K, P = float(pop_prev), float(samp_prev)
    zv = norm.pdf(norm.ppf(K))
    fac = K ** 2 * (1 - K) ** 2 / P / (1 - P) / zv ** 2
    h2_liab = h2 * fac
    var = (se * fac) ** 2
    return h2_liab, var ** 0.5, chi2.sf(h2_liab ** 2 / var, 1)


def write_trace_files(trace_sums, M_table, *, pheno_file, trace_dir, num_indv, num_snp, num_jack, num_bin,
                      num_random_vec) -> str:
    """`run_<pheno>.tr` / `.MN` in the SUMRHE trace-summary format -- base.py:831-855."""
    name = f"run_{os.path.basename(pheno_file) if pheno_file is not None else None}"
    prefix = os.path.join(trace_dir, name) if (trace_dir and os.path.isdir(trace_dir)) else name
    with open(prefix + ".MN", "w") as fd:
        fd.write("NSAMPLE,NSNPS,NBLKS,NBINS,K\n")
        fd.write(f"{num_indv:.0f},{num_snp:.0f},{num_jack:.0f},{num_bin:.0f},{num_random_vec:.0f}")
    with open(prefix + ".tr", "w") as fd:
        fd.write(",".join(f"LD_SUM_{i:d}" for i in range(num_bin)) + ",NSNPS_JACKKNIFE\n")
        for j in range(num_jack + 1):
            for k in range(num_bin):
                cells = ",".join(f"{trace_sums[j, k, l]:.3f}" for l in range(num_bin))
                fd.write(f"{cells},{M_table[j, k]:.0f}\n")
    return prefix
