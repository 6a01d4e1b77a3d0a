"""CPU ORACLE for the PyRHE trace-estimation hot path.  TEST INFRASTRUCTURE ONLY.

This is a restatement, in numpy + torch-CPU fp32, of the algorithm the reference
runs per jackknife block.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import it; the product (`pyrhe_b200/`) never does and has no CPU
fallback.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks this file against
`tests/golden/*.npz`, which `tools/make_golden.py` produced by running the
unmodified reference classes from /root/reference in this container (with the
`bed_reader` test shim, integer seed, OMP_NUM_THREADS=1).  The arithmetic keeps
the reference's dtype flow: float32 genotypes and products via
torch, float64 state from `aggregate` on.

Third-party arithmetic restated here: `bed_reader==1.0.2` (pyrhe/setup.py:13)
decodes PLINK-1 SNP-major `.bed` with count_A1=True, float32, NaN = missing;
the reference then swaps 0<->2 (base.py:352-355).  Net map of the 2-bit code:
00 -> 0, 10 -> 1, 11 -> 2, 01 -> NaN (count of the .bim A2 allele).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import scipy.linalg
import torch

_CODE_TO_A2 = np.array([0.0, np.nan, 1.0, 2.0], dtype=np.float32)

# Working precision of the block products.  float32 is the reference's (mat_mul.py:12 casts every operand);
# `run(problem, f64=True)` repeats the same algorithm in float64 to measure how far the reference's own fp32
# arithmetic sits from the exact answer (the yardstick for the parity tolerances).
_DT = {"np": np.float32, "torch": torch.float32}


# --------------------------------------------------------------------------- decode
def decode_bed_rows(packed: np.ndarray, n_indv: int) -> np.ndarray:
    """packed [m, ceil(N0/4)] uint8 -> float32 [N0, m] (F-order), A2 count, NaN = missing.

    bed_reader.open_bed.read(index=np.s_[::1, a:b]) as called at base.py:341 plus the
    0<->2 swap of base.py:352-355.
    """
    m = packed.shape[0]
    codes = np.empty((m, packed.shape[1] * 4), dtype=np.uint8)
    for shift in range(4):
        codes[:, shift::4] = (packed >> (2 * shift)) & 3
    return np.asfortranarray(_CODE_TO_A2.astype(_DT["np"])[codes[:, :n_indv]].T)


def block_range(M: int, J: int, j: int):
    """base.py:362-371 -- equal steps, the last block takes the remainder."""
    step = M // J
    start = j * step
    return start, start + (step if j < J - 1 else step + M % J)


# --------------------------------------------------------------------------- imputation
def _fill_from_uniform(p, rval):
    """base.py:265-274 with p = observed_mean * 0.5 (a numpy float32 scalar)."""
    d0 = (1 - p) * (1 - p)
    d1 = 2 * p * (1 - p)
    if rval < d0:
        return 0
    if rval < d0 + d1:
        return 1
    return 2


def impute_block(X: np.ndarray, method: str, rng=np.random) -> np.ndarray:
    """base.py:277-289.  One uniform draw PER SNP in binary mode (drawn even with no
    missing entry, Q5); "mean" fills 0 BEFORE standardisation (Q4).  In place."""
    for s in range(X.shape[1]):
        col = X[:, s]
        miss = np.isnan(col)
        if method == "binary":
            X[miss, s] = _fill_from_uniform(np.nanmean(col) * 0.5, rng.random())
        else:
            X[miss, s] = 0
    return X


# --------------------------------------------------------------------------- fp32 operator boundary
def _t32(a):
    """mat_mul.py:4-15 -- numpy -> torch float32 (strides kept: they select the BLAS path)."""
    if isinstance(a, torch.Tensor):
        return a
    a = np.asarray(a)
    if not a.flags.writeable:
        a = a.copy(order="K")
    return torch.from_numpy(a).to(_DT["torch"])


def mm(*mats) -> np.ndarray:
    """mat_mul.py:17-31 -- chained fp32 `@`, result back in numpy float32."""
    r = _t32(mats[0])
    for m_ in mats[1:]:
        r = r @ _t32(m_)
    return r.numpy()


def standardize(G: np.ndarray) -> np.ndarray:
    """base.py:291-296 -- binomial-variance standardisation, float32 throughout."""
    mu = np.mean(G, axis=0)
    return (G - mu) * (1 / np.sqrt(mu * (1 - 0.5 * mu)))


def dominance_standardized(G: np.ndarray) -> np.ndarray:
    """rhe_dom.py:15-41 -- dominance coding then scaling by 1/(2 maf (1-maf))."""
    maf = np.mean(G, axis=0) / 2
    enc = np.zeros_like(G, dtype=np.float64)
    enc += (G == 1) * (2 * maf[np.newaxis, :])
    enc += (G == 2) * (4 * maf[np.newaxis, :] - 2)
    return (enc - np.mean(enc, axis=0)) * (1 / (2 * maf * (1 - maf)))


# --------------------------------------------------------------------------- problem description
@dataclass
class OracleProblem:
    packed: np.ndarray            # [M, ceil(N0/4)] uint8, the .bed payload (no magic)
    n_indv_original: int
    annot: np.ndarray             # [M, K] 0/1
    Z: np.ndarray                 # [N, B] float64 (all_zb, base.py:176)
    y: np.ndarray                 # [N, 1] float64, centred (base.py:154)
    num_jack: int
    W: Optional[np.ndarray] = None   # [N, C] covariates (after row filtering)
    missing_indv: tuple = ()
    impute: str = "binary"
    seed: int = 0
    model: str = "rhe"            # rhe | rhe_dom | genie
    genie_model: str = "G+GxE+NxE"
    env: Optional[np.ndarray] = None  # [N] (one environment, file_processing.py:212-220)
    _derived: dict = field(default_factory=dict)


class Oracle:
    """Non-streaming reference flow: pre_compute -> aggregate -> (T_j, q_j) -> solve."""

    def __init__(self, p: OracleProblem):
        self.p = p
        self.N = p.Z.shape[0]
        self.B = p.Z.shape[1]
        self.M_snps, self.K = p.annot.shape
        self.J = p.num_jack
        self.use_cov = p.W is not None
        self.len_bin = (p.annot == 1).sum(0)
        if self.use_cov:
            self.Q = np.linalg.pinv(p.W.T @ p.W)                      # base.py:151
            self.UZ = p.W @ self.Q @ (p.W.T @ p.Z)                    # base.py:178
        self.n_env = 0
        self.n_gxe = 0
        if p.model == "rhe":
            self.E = self.K
            m_last = self.len_bin
        elif p.model == "rhe_dom":
            self.E = 2 * self.K                                        # rhe_dom.py:7-13
            m_last = np.concatenate([self.len_bin, self.len_bin])
        elif p.model == "genie":
            self.n_env = 1
            self.n_gxe = self.K
            if p.genie_model == "G":                                   # genie.py:26-44
                self.E = self.K
                m_last = self.len_bin
            elif p.genie_model == "G+GxE+NxE":
                self.E = 2 * self.K + 1
                m_last = np.concatenate([self.len_bin, self.len_bin, [1]])
            else:
                raise NotImplementedError("reference 'G+GxE' mislabels the last GxE row as NxE "
                                          "(SURVEY.md §9.3 Q7); it has no parity target")
        else:
            raise ValueError(p.model)
        sh = (self.E, self.J + 1, self.B, self.N)                       # base.py:419-429
        self.XXz = np.zeros(sh)
        self.yXXy = np.zeros((self.E, self.J + 1))
        self.Mjk = np.zeros((self.J + 1, self.E), dtype=np.int64)
        self.Mjk[self.J] = m_last
        if self.use_cov:
            self.UXXz = np.zeros(sh)
            self.XXUz = np.zeros(sh)

    # ---- per-statistic helpers (base.py:396-417)
    def _y_res(self):
        if not self.use_cov:
            return self.p.y
        W = self.p.W
        Q = np.linalg.pinv(W.T @ W)
        return self.p.y - W @ (Q @ (W.T @ self.p.y))

    def _xxz(self, X, vec):
        return mm(X, mm(X.T, vec.reshape(-1, 1))).flatten()

    def _uxxz(self, v):
        return mm(self.p.W, mm(self.Q, mm(self.p.W.T, v))).flatten()

    def _yxxy(self, X):
        v = mm(X.T, self._y_res())
        return mm(v.T, v)[0][0]

    def _fill(self, e, j, X):
        """One (estimate, block) cell: rhe.py:13-22 inner loop."""
        for b in range(self.B):
            self.XXz[e, j, b, :] = self._xxz(X, self.p.Z[:, b])
            if self.use_cov:
                self.UXXz[e, j, b, :] = self._uxxz(self.XXz[e][j][b])
                self.XXUz[e, j, b, :] = self._xxz(X, self.UZ[:, b])
        self.yXXy[e][j] = self._yxxy(X)

    # ---- a1-a3: block read / impute / bin gather
    def block_bins(self, j, return_block=False):
        p = self.p
        start, end = block_range(self.M_snps, self.J, j)
        sub = decode_bed_rows(p.packed[start:end], p.n_indv_original)
        if len(p.missing_indv):
            sub = np.delete(sub, list(p.missing_indv), axis=0)          # base.py:344
        sub = sub.copy()                                                # base.py:352 (C-order copy)
        np.random.seed(p.seed)                                          # base.py:510
        sub = impute_block(sub, p.impute)
        if return_block:
            return sub
        ann = p.annot[start:end]
        return [sub[:, np.nonzero(ann[:, k])[0].tolist()] for k in range(self.K)]  # base.py:315-336

    # ---- a11: model hooks
    def pre_compute(self):
        p, K = self.p, self.K
        for j in range(self.J):
            bins = self.block_bins(j)
            for k, G in enumerate(bins):
                X = standardize(G)
                self.Mjk[j][k] = self.Mjk[self.J][k] - X.shape[1]
                self._fill(k, j, X)
                if p.model == "rhe_dom":                                # rhe_dom.py:57-68
                    Xd = dominance_standardized(G)
                    self.Mjk[j][k + K] = self.Mjk[self.J][k + K] - Xd.shape[1]
                    self._fill(k + K, j, Xd)
            if p.model == "genie" and p.genie_model != "G":             # genie.py:61-82
                envc = p.env.reshape(-1, 1)
                for k, G in enumerate(bins):
                    X = standardize(G)
                    kg = k + K                                          # (e+1)*k + K with e = 0
                    self.Mjk[j][kg] = self.Mjk[self.J][kg] - X.shape[1]
                    Xg = (_t32(X) * _t32(envc)).numpy()                 # elem_mul, mat_mul.py:34-48
                    self._fill(kg, j, Xg)
                self.Mjk[j][2 * K] = 1
        self.aggregate()

    # ---- a13: totals + leave-one-out by subtraction (base.py:465-500)
    def aggregate(self):
        J, B = self.J, self.B
        arrays = [self.XXz] + ([self.UXXz, self.XXUz] if self.use_cov else [])
        for e in range(self.E):
            if e < self.E - (self.n_env if self.p.genie_model == "G+GxE+NxE" and self.p.model == "genie" else 0):
                for j in range(J):
                    for A in arrays:
                        A[e, J] += A[e, j]
                    self.yXXy[e][J] += self.yXXy[e][j]
            else:
                # NxE row: X = diag(env).  The reference multiplies by a dense N x N matrix
                # (base.py:474); with X diagonal each fp32 product reduces to env*(env*z).
                env32 = self.p.env.astype(_DT["np"])
                for b in range(B):
                    z32 = self.p.Z[:, b].astype(_DT["np"])
                    self.XXz[e][J][b] = env32 * (env32 * z32)
                v = (env32.reshape(-1, 1) * self._y_res().astype(_DT["np"]))
                self.yXXy[e][J] = mm(v.T, v)[0][0]
                if self.use_cov:                                        # Q8: only b = B-1 is filled
                    b = B - 1
                    self.UXXz[e][J][b] = self._uxxz(self.XXz[e][J][b])
                    u32 = self.UZ[:, b].astype(_DT["np"])
                    self.XXUz[e][J][b] = env32 * (env32 * u32)
            for j in range(J):
                for A in arrays:
                    A[e, j] = A[e, J] - A[e, j]
                self.yXXy[e][j] = self.yXXy[e][J] - self.yXXy[e][j]

    # ---- a14: normal equations (base.py:568-628; genie.py:84-94)
    def lhs_rhs(self, j):
        E, B, N = self.E, self.B, self.N
        T = np.zeros((E + 1, E + 1))
        q = np.zeros((E + 1, 1))
        W = self.p.W
        for a in range(E):
            for c in range(E):
                Ma, Mc = self.Mjk[j][a], self.Mjk[j][c]
                B1, B2 = self.XXz[a][j], self.XXz[c][j]
                T[a, c] += np.sum(B1 * B2)
                if self.use_cov:
                    h3 = W @ (self.Q @ (W.T @ B1.T))
                    r1 = np.sum(h3.T * B2)
                    r2 = np.sum(self.XXUz[a][j] * self.UXXz[c][j])
                    T[a, c] += (r2 - 2 * r1)
                T[a, c] /= B
                T[a, c] = T[a, c] / (Ma * Mc) if (Ma * Mc) != 0 else 0
        for a in range(E):
            Ma = self.Mjk[j][a]
            if self.p.model == "genie" and a >= self.K:
                tr = np.sum(self.XXz[a][j] * self.p.Z.T) / (B * Ma)
            else:
                tr = N
            if self.use_cov:
                tr = tr - 1 / (B * Ma) * np.sum(self.XXz[a][j] * self.UZ.T)
            T[a, E] = T[E, a] = tr
            q[a] = self.yXXy[(a, j)] / Ma if Ma != 0 else 0
        T[E, E] = N if not self.use_cov else N - W.shape[1]
        yr = self._y_res()
        q[E] = yr.T @ yr
        return T, q

    def trace_sum(self, T, j):
        """base.py:598-599,827-829 -- LD_SUM entries of the .tr file."""
        E = self.E
        out = np.zeros((E, E))
        for a in range(E):
            for c in range(E):
                Ma, Mc = self.Mjk[j][a], self.Mjk[j][c]
                out[a, c] = (T[a, c] - self.N) * (Ma * Mc) / pow(self.N, 2) if Ma * Mc != 0 else 0
        return out

    def estimate(self):
        """base.py:630-678 with method="QR" (base.py:306-312, default of __call__ :874)."""
        Ts, qs, sig = [], [], []
        for j in range(self.J + 1):
            T, q = self.lhs_rhs(1 if (self.J == 1 and j == 0) else j)
            Qm, R = scipy.linalg.qr(T)
            sig.append(np.ravel(scipy.linalg.solve_triangular(R, np.dot(Qm.T, q))))
            Ts.append(T)
            qs.append(q.ravel())
        return np.array(Ts), np.array(qs), np.array(sig)


def jackknife_se(ests: np.ndarray, J: int) -> np.ndarray:
    """base.py:680-703."""
    mean = ests.mean(axis=0)
    return np.sqrt((J - 1) * ((ests - mean) ** 2).sum(axis=0) / J)


def run(problem: OracleProblem, f64: bool = False) -> dict:
    """`f64=True`: the same algorithm with float64 block products (not the reference's arithmetic: a yardstick)."""
    saved = dict(_DT)
    if f64:
        _DT.update(np=np.float64, torch=torch.float64)
    try:
        o = Oracle(problem)
        o.pre_compute()
        T, q, sig = o.estimate()
    finally:
        _DT.update(saved)
    out = dict(T=T, q=q, sigma_jack=sig[:-1], sigma_total=sig[-1],
               sigma_se=jackknife_se(sig[:-1], o.J), M=o.Mjk, XXz=o.XXz, yXXy=o.yXXy)
    if o.use_cov:
        out.update(UXXz=o.UXXz, XXUz=o.XXUz)
    return out
