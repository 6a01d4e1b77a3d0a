"""Small host-side (numpy fp64) pieces of the path: block partition, right-hand sides,
covariate projections, the imputation rule.  O(N (B + C)) work at most; the O(N M)
work lives in the CUDA library.
"""
from __future__ import annotations

import numpy as np

from .assemble import HostTerms, PathPlan


def block_ranges(M: int, J: int):
    """SNP range of every jackknife block: equal steps of M // J, the last block takes
    the remainder (/root/reference/pyrhe/src/base/base.py:362-371)."""
    step = M // J
    return [(j * step, (j + 1) * step if j < J - 1 else M) for j in range(J)]


def impute_uniforms(seed: int, count: int) -> np.ndarray:
    """The uniforms `np.random.random()` yields right after `np.random.seed(seed)`.

    The non-streaming reference re-seeds at every block (base.py:510) and draws one
    uniform per SNP of the block, even when nothing is missing (base.py:281-285), so
    block j uses the first m_j values of this one stream."""
    return np.random.RandomState(seed).random_sample(count)


def binary_fill_values(n1, n2, n_miss, n_kept, seed) -> np.ndarray:
    """Fill value (0/1/2) per SNP of one block under "binary" imputation.

    Mirrors the float32 arithmetic of base.py:265-285: p = float32(nanmean) * 0.5,
    thresholds in float32, and numpy compares the float64 draw AFTER casting it to
    float32 (weak-scalar promotion).  The CUDA library implements the same rule with
    round-to-nearest intrinsics; this function is its specification."""
    n1 = np.asarray(n1, dtype=np.float64)
    n2 = np.asarray(n2, dtype=np.float64)
    observed = n_kept - np.asarray(n_miss, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean32 = ((n1 + 2 * n2) / observed).astype(np.float32)
    p = mean32 * np.float32(0.5)
    one = np.float32(1)
    d0 = (one - p) * (one - p)
    d1 = (np.float32(2) * p) * (one - p)
    u32 = impute_uniforms(seed, len(n1)).astype(np.float32)
    return np.where(u32 < d0, 0, np.where(u32 < d0 + d1, 1, 2)).astype(np.int64)


def host_terms(plan: PathPlan, Z: np.ndarray, W, Y: np.ndarray, env) -> tuple:
    """(HostTerms, Y_res).  Z [N,B], W [N,C] or None, Y [N,Ty] centred, env [N] or None."""
    N = Z.shape[0]
    ht = HostTerms(N=N)
    if W is not None:
        ht.Q = np.linalg.pinv(W.T @ W)                         # base.py:151
        ht.WtZ = W.T @ Z
        Y_res = Y - W @ (ht.Q @ (W.T @ Y))                     # base.py:396-401
    else:
        Y_res = Y
    ht.yy_res = np.einsum("nt,nt->t", Y_res, Y_res)            # base.py:625-626
    if plan.has_nxe:
        e2 = np.asarray(env, dtype=np.float64) ** 2
        ht.nxe_tr = float(np.sum((e2[:, None] * Z) * Z))
        ht.nxe_yxxy = np.einsum("nt,nt->t", e2[:, None] * Y_res, Y_res)
        if W is not None:
            UZ = W @ (ht.Q @ ht.WtZ)                           # base.py:178
            ht.nxe_H = W.T @ (e2[:, None] * Z)
            ht.nxe_WtLU = W.T @ (e2[:, None] * UZ)
    return ht, Y_res


def rhs_matrix(plan: PathPlan, Z, W, Y_res, env, keep: np.ndarray, dtype=np.float64, width: int | None = None):
    """Right-hand sides in FILE order (row i = individual i of the .fam; rows of dropped
    individuals are zero so they add nothing to X^T R).

    Returns (R [n_sets*Rs, width], rowscale [n_sets, width]), `width` >= N0 (default N0; the engine asks for its padded
    row length and fp32 so that no second pass over the matrix is needed).  Set 0 = [Z | W | y_res]; set 1 (GxE) =
    env * set 0, because (diag(env) X)^T r = X^T (env * r) (genie.py:65-67).  `rowscale` is
    what pass B multiplies its output rows by: keep mask, times env for the GxE set."""
    n0 = keep.shape[0]
    width = n0 if width is None else int(width)
    assert width >= n0
    cols = [np.asarray(c, dtype=np.float64) for c in [Z] + ([W] if W is not None else []) + [Y_res]]
    R = np.zeros((plan.n_sets * plan.Rs, width), dtype=dtype)
    rowscale = np.zeros((plan.n_sets, width), dtype=dtype)
    # column block by column block, straight into its rows (no [N, Rs] concatenation, no boolean-mask scatter: at
    # N = 500k those cost 70-160 ms of every set_rhs; values are the same to the last bit)
    idx = None if keep.all() else np.flatnonzero(keep)

    def put(row0, block):
        c = block.shape[1]
        if idx is None:
            for a in range(0, n0, 8192):                       # cache-sized pieces of the transposition
                R[row0: row0 + c, a: min(a + 8192, n0)] = block[a: a + 8192].T
        else:
            R[row0: row0 + c, idx] = block.T
        return row0 + c

    r = 0
    for c in cols:
        r = put(r, c)
    assert r == plan.Rs
    rowscale[0, :n0][keep] = 1.0
    if plan.n_sets == 2:
        e = np.asarray(env, dtype=np.float64)
        for c in cols:
            r = put(r, c * e[:, None])
        rowscale[1, :n0][keep] = e
    return R, rowscale
