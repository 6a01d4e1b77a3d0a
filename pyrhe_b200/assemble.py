"""Host-side (fp64 numpy) assembly of the per-jackknife normal equations T sigma = q.

The reference builds T and q from N-length state arrays
(/root/reference/pyrhe/src/base/base.py:568-628, genie.py:84-94).  Here the device
returns only small "pieces" per jackknife block j and estimate e:

* ``XX[j][a, c] = <L_a, L_c>`` -- Gram of the leave-one-out vectors
  ``L_e = S_e - P_ej`` (``XXz`` in the reference), summed over B x N;
* ``G[j][e]`` -- the ``Rs x Rs`` Gram of the standardised pass-A products
  ``t_s = X_s^T [Z | W | y]`` over the SNPs of the (block, bin):
  ``G = sum_s t_s t_s^T``.  Because ``X X^T`` is symmetric every covariate term of
  base.py:583-593,612-618 is a contraction of its sub-blocks:
  ``W^T XXz = G[W, Z]``, ``W^T X X^T W = G[W, W]``, ``yXXy = G[y, y]``,
  ``<XXz, Z> = tr G[Z, Z]``.

Everything is linear in ``X X^T`` so leave-one-out is total minus block.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np


@dataclass
class PathPlan:
    """Column / estimate layout shared by the host code and the CUDA library."""
    model: str                 # "rhe" | "rhe_dom" | "genie"
    K: int                     # bins
    B: int                     # random vectors
    C: int                     # covariates (0 = none)
    Ty: int = 1                # phenotype columns carried through pass A
    genie_model: str = "G+GxE+NxE"

    @property
    def n_ops(self):           # genotype operands: additive (+ dominance)
        return 2 if self.model == "rhe_dom" else 1

    @property
    def n_sets(self):          # right-hand-side sets: plain (+ env-scaled for GxE)
        return 2 if (self.model == "genie" and self.genie_model != "G") else 1

    @property
    def has_nxe(self):
        return self.model == "genie" and self.genie_model == "G+GxE+NxE"

    @property
    def n_groups(self):
        return self.n_ops * self.n_sets

    @property
    def Rs(self):              # columns per set: [Z (B) | W (C) | y (Ty)]
        return self.B + self.C + self.Ty

    @property
    def E_reg(self):           # estimates that come from genotype blocks
        return self.n_groups * self.K

    @property
    def E(self):
        return self.E_reg + (1 if self.has_nxe else 0)

    def cols_Z(self):
        return slice(0, self.B)

    def cols_W(self):
        return slice(self.B, self.B + self.C)

    def col_y(self, trait):
        return self.B + self.C + trait


@dataclass
class HostTerms:
    """Small fp64 quantities the host computes once (never N x m work)."""
    N: int
    Q: Optional[np.ndarray] = None        # (W^T W)^+            base.py:151
    WtZ: Optional[np.ndarray] = None      # W^T Z   [C, B]
    yy_res: np.ndarray = field(default_factory=lambda: np.zeros(1))  # ||y_res||^2 per trait  base.py:625-626
    # NxE row (X = diag(env), base.py:472-481): all closed-form, O(N B)
    nxe_H: Optional[np.ndarray] = None    # W^T (env^2 * Z)      [C, B]
    nxe_WtLU: Optional[np.ndarray] = None  # W^T (env^2 * UZ)     [C, B]
    nxe_tr: float = 0.0                   # <env^2 * Z, Z>
    nxe_yxxy: Optional[np.ndarray] = None  # sum (env * y_res)^2  per trait


def _last_col_only(A):
    out = np.zeros_like(A)
    out[:, -1] = A[:, -1]
    return out


def normal_equations(plan: PathPlan, ht: HostTerms, XX: np.ndarray, G_loo: np.ndarray, M_row: np.ndarray,
                     trait: int = 0, nxe_quirk: bool = True):
    """(T, q) for one jackknife sample.

    XX     [E, E]            Gram of the leave-one-out XXz vectors (includes the NxE row)
    G_loo  [E_reg, Rs, Rs]   leave-one-out pass-A Gram per estimate
    M_row  [E]               SNP counts of this jackknife sample (base.py:576-577)
    nxe_quirk: reproduce base.py:479-481 (UXXz/XXUz of the NxE row filled for b = B-1 only).
    """
    E, B, C, N = plan.E, plan.B, plan.C, ht.N
    use_cov = C > 0
    T = np.zeros((E + 1, E + 1))
    q = np.zeros((E + 1, 1))
    zs, ws, yc = plan.cols_Z(), plan.cols_W(), plan.col_y(trait)

    if use_cov:
        QWtZ = ht.Q @ ht.WtZ
        H, Hq, WtLU, full_b = [], [], [], []
        for e in range(E):
            if e < plan.E_reg:
                H_e = G_loo[e][ws, zs]                       # W^T XXz_e
                WtLU_e = G_loo[e][ws, ws] @ QWtZ             # W^T XXUz_e = (W^T X X^T W) Q W^T Z
                full = True
            else:
                H_e, WtLU_e = ht.nxe_H, ht.nxe_WtLU
                full = not nxe_quirk
            H.append(H_e)
            # the arrays that enter the <XXUz, UXXz> term (zero except b = B-1 for the NxE row)
            Hq.append(H_e if full else _last_col_only(H_e))
            WtLU.append(WtLU_e if full else _last_col_only(WtLU_e))

    for a in range(E):
        for c in range(E):
            Ma, Mc = M_row[a], M_row[c]
            v = XX[a, c]
            if use_cov:
                r1 = np.sum(H[a] * (ht.Q @ H[c]))             # <U XXz_a, XXz_c>      base.py:584-587
                r2 = np.sum(WtLU[a] * (ht.Q @ Hq[c]))         # <XXUz_a, UXXz_c>      base.py:589-591
                v += r2 - 2 * r1
            v /= B
            T[a, c] = v / (Ma * Mc) if (Ma * Mc) != 0 else 0

    for a in range(E):
        Ma = M_row[a]
        if plan.model == "genie" and a >= plan.K:             # genie.py:84-94 (Hutchinson trace)
            zz = np.trace(G_loo[a][zs, zs]) if a < plan.E_reg else ht.nxe_tr
            tr = zz / (B * Ma)
        else:
            tr = N                                            # rhe.py:24-26
        if use_cov:
            tr = tr - np.sum(H[a] * QWtZ) / (B * Ma)          # base.py:612-618
        T[a, E] = T[E, a] = tr
        yxxy = G_loo[a][yc, yc] if a < plan.E_reg else ht.nxe_yxxy[trait]
        q[a] = yxxy / Ma if Ma != 0 else 0
    T[E, E] = N - C                                           # base.py:623
    q[E] = ht.yy_res[trait]
    return T, q


def trace_sums_row(T: np.ndarray, M_row: np.ndarray, N: int, E: int) -> np.ndarray:
    """LD_SUM[k, l] = (T[k,l] - N) M_k M_l / N^2 (base.py:598-599, 827-829)."""
    Mk = np.asarray(M_row, dtype=np.float64)
    MM = np.outer(Mk, Mk)
    out = (T[:E, :E] - N) * MM / float(N) ** 2
    out[MM == 0] = 0
    return out


def loo_grams(G_blk: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
    """[J, ...] per-block pieces -> [J + 1, ...]: total minus block j for j < J, the total in slot J.  `out` lets a
    caller that runs this every step reuse one buffer (a fresh 1.6 MB array costs more than the arithmetic)."""
    J = G_blk.shape[0]
    if out is None or out.shape != (J + 1,) + G_blk.shape[1:]:
        out = np.empty((J + 1,) + G_blk.shape[1:], dtype=np.float64)
    np.sum(G_blk, axis=0, out=out[J])
    np.subtract(out[J][None], G_blk, out=out[:J])
    return out


def normal_equations_prepare(plan: PathPlan, ht: HostTerms, G_loo: np.ndarray, M_tab: np.ndarray,
                             nxe_quirk: bool = True) -> dict:
    """The part of `normal_equations_batch` that needs only the per-bin Gram pieces (every covariate term, the trace
    column, yXXy): `RheEngine.run(gram_hook=...)` calls it on the host while the device still forms the leave-one-out
    Grams `XX`, so only `normal_equations_finish` is left once `XX` arrives."""
    E, E_reg, B, C, N = plan.E, plan.E_reg, plan.B, plan.C, ht.N
    S = G_loo.shape[0]
    zs, ws = plan.cols_Z(), plan.cols_W()
    Mf = np.asarray(M_tab, dtype=np.float64)
    MM = Mf[:, :, None] * Mf[:, None, :]
    corr = None
    H = QWtZ = None
    if C > 0:
        QWtZ = ht.Q @ ht.WtZ
        H = np.empty((S, E, C, B))
        WtLU = np.empty((S, E, C, B))
        H[:, :E_reg] = G_loo[:, :, ws, zs]
        WtLU[:, :E_reg] = G_loo[:, :, ws, ws] @ QWtZ
        Hq = H
        if plan.has_nxe:
            H[:, E_reg] = ht.nxe_H
            WtLU[:, E_reg] = ht.nxe_WtLU
            if nxe_quirk:
                Hq = H.copy()
                Hq[:, E_reg, :, :-1] = 0
                WtLU[:, E_reg, :, :-1] = 0
        # batched matmuls (the einsum forms of the same contractions are ~6x slower in numpy)
        QH = ht.Q @ H                                                   # [S, E, C, B]
        QHq = QH if Hq is H else ht.Q @ Hq
        flat = lambda a: a.reshape(S, E, C * B)
        r1 = flat(H) @ flat(QH).transpose(0, 2, 1)                      # sum_cb H[s,a,c,b] QH[s,e,c,b]
        r2 = flat(WtLU) @ flat(QHq).transpose(0, 2, 1)
        corr = r2 - 2 * r1
    tr = np.full((S, E), float(N))
    if plan.model == "genie" and E > plan.K:
        zz = np.einsum("sebb->se", G_loo[:, plan.K:E_reg, zs, zs])
        tr[:, plan.K:E_reg] = zz / (B * Mf[:, plan.K:E_reg])
        if plan.has_nxe:
            tr[:, E_reg] = ht.nxe_tr / (B * Mf[:, E_reg])
    if C > 0:
        tr = tr - np.einsum("secb,cb->se", H, QWtZ) / (B * Mf)
    return dict(plan=plan, ht=ht, G_loo=G_loo, Mf=Mf, MM=MM, corr=corr, tr=tr)


def normal_equations_finish(prep: dict, XX: np.ndarray, trait: int = 0):
    """`normal_equations_prepare` + the leave-one-out Grams XX [S, E, E]  ->  T [S, E+1, E+1], q [S, E+1]."""
    plan, ht, G_loo, Mf, MM = prep["plan"], prep["ht"], prep["G_loo"], prep["Mf"], prep["MM"]
    E, E_reg, B, C, N = plan.E, plan.E_reg, plan.B, plan.C, ht.N
    S = XX.shape[0]
    yc = plan.col_y(trait)
    V = XX.astype(np.float64).copy()
    T = np.zeros((S, E + 1, E + 1))
    q = np.zeros((S, E + 1))
    if prep["corr"] is not None:
        V += prep["corr"]
    V /= B
    np.divide(V, MM, out=T[:, :E, :E], where=MM != 0)
    T[:, :E, E] = prep["tr"]
    T[:, E, :E] = prep["tr"]
    T[:, E, E] = N - C
    yxxy = np.empty((S, E))
    yxxy[:, :E_reg] = G_loo[:, :, yc, yc]
    if plan.has_nxe:
        yxxy[:, E_reg] = ht.nxe_yxxy[trait]
    np.divide(yxxy, Mf, out=q[:, :E], where=Mf != 0)
    q[:, E] = ht.yy_res[trait]
    return T, q


def normal_equations_batch(plan: PathPlan, ht: HostTerms, XX: np.ndarray, G_loo: np.ndarray, M_tab: np.ndarray,
                           trait: int = 0, nxe_quirk: bool = True):
    """Vectorised `normal_equations` over all jackknife samples (the production path).

    XX [S, E, E], G_loo [S, E_reg, Rs, Rs], M_tab [S, E]  ->  T [S, E+1, E+1], q [S, E+1]."""
    return normal_equations_finish(normal_equations_prepare(plan, ht, G_loo, M_tab, nxe_quirk), XX, trait)
