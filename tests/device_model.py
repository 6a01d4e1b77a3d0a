"""TEST INFRASTRUCTURE: fp64 numpy model of the algorithm the CUDA library runs.

It follows DESIGN.md §3 step by step (packed codes -> masked popcounts -> fill ->
raw products G^T R -> rank-1 standardisation fix-up -> per-bin Gram + pass-B
weights -> P = X (X^T Z)), so the algebraic restructuring and the host assembly in
pyrhe_b200/assemble.py can be checked against the golden vectors on a box with no
GPU.  The product never imports this file.
"""
import numpy as np

from pyrhe_b200.assemble import PathPlan, HostTerms, normal_equations
from pyrhe_b200.hostmath import binary_fill_values, block_ranges, host_terms, rhs_matrix


def decode_codes(packed, n0):
    m = packed.shape[0]
    codes = np.empty((m, packed.shape[1] * 4), dtype=np.uint8)
    for sh in range(4):
        codes[:, sh::4] = (packed >> (2 * sh)) & 3
    return codes[:, :n0]


def run_model(packed, n0, annot, Z, Y_res, W, env, missing_indv, J, impute, seed, plan: PathPlan):
    """Returns dict(XX [J+1,E,E], G_blk [J,E_reg,Rs,Rs], M [J+1,E], P, S)."""
    M_snps, K = annot.shape
    keep = np.ones(n0, dtype=bool)
    keep[list(missing_indv)] = False
    N = int(keep.sum())
    R, rowscale = rhs_matrix(plan, Z, W, Y_res, env, keep)        # [n_sets*Rs, n0], [n_sets, n0]
    colsum = R.sum(axis=1)
    Rs, B, E, E_reg = plan.Rs, plan.B, plan.E, plan.E_reg
    P = np.zeros((J, E, B, n0))
    G_blk = np.zeros((J, E_reg, Rs, Rs))
    Mjk = np.zeros((J + 1, E), dtype=np.int64)
    len_bin = (annot == 1).sum(0)
    Mjk[J, :E_reg] = np.tile(len_bin, plan.n_groups)
    if plan.has_nxe:
        Mjk[:, E_reg] = 1
    val_of_code = np.array([0, 0, 1, 2])
    for j, (a, b) in enumerate(block_ranges(M_snps, J)):
        codes = decode_codes(packed[a:b], n0)
        ck = codes[:, keep]
        n1 = (ck == 2).sum(1)
        n2 = (ck == 3).sum(1)
        nm = (ck == 1).sum(1)
        fill = binary_fill_values(n1, n2, nm, N, seed) if impute == "binary" else np.zeros(b - a, np.int64)
        g = val_of_code[codes] + (codes == 1) * fill[:, None]    # [m, n0] imputed A2 counts
        n1e = n1 + nm * (fill == 1)
        n2e = n2 + nm * (fill == 2)
        mu = (n1e + 2 * n2e) / N
        var = mu * (1 - mu / 2)
        T_add = g @ R.T                                          # raw pass-A products
        T_i2 = (g == 2).astype(np.float64) @ R.T
        sub_annot = annot[a:b]
        for grp in range(plan.n_groups):
            op, st = (grp, 0) if plan.n_ops == 2 else (0, grp)
            cols = slice(st * Rs, (st + 1) * Rs)
            if op == 0:                                          # additive: x = (g - mu) r
                r = 1 / np.sqrt(var)
                t = r[:, None] * (T_add[:, cols] - mu[:, None] * colsum[cols])
                w1 = (r[:, None] * t[:, :B])                     # weight of [g == 1]
                w2 = 2 * w1                                      # weight of [g == 2]
                shift = mu[:, None] * w1
            else:                                                # dominance: h = mu g - 2 [g == 2]
                r = 1 / var
                eta = mu * mu - 2 * n2e / N
                t = r[:, None] * (mu[:, None] * T_add[:, cols] - 2 * T_i2[:, cols] - eta[:, None] * colsum[cols])
                u = r[:, None] * t[:, :B]
                w1 = mu[:, None] * u
                w2 = 2 * w1 - 2 * u
                shift = eta[:, None] * u
            for k in range(K):
                rows = np.nonzero(sub_annot[:, k])[0]
                e = grp * K + k
                Mjk[j, e] = Mjk[J, e] - rows.size
                G_blk[j, e] = t[rows].T @ t[rows]
                i1 = (g[rows] == 1).astype(np.float64)
                i2 = (g[rows] == 2).astype(np.float64)
                acc = w1[rows].T @ i1 + w2[rows].T @ i2 - shift[rows].sum(0)[:, None]   # [B, n0]
                P[j, e] = acc * rowscale[st]
    S = P.sum(axis=0)
    if plan.has_nxe:
        e2 = np.zeros(n0)
        e2[keep] = env.astype(np.float64) ** 2
        Zfull = np.zeros((n0, B))
        Zfull[keep] = Z
        S[E_reg] = (e2[:, None] * Zfull).T
    XX = np.zeros((J + 1, E, E))
    for j in range(J + 1):
        L = (S - P[j]) if j < J else S
        Lf = L.reshape(E, -1)
        XX[j] = Lf @ Lf.T
    return dict(XX=XX, G_blk=G_blk, M=Mjk, P=P, S=S, N=N)


def assemble_all(plan, ht: HostTerms, pieces, J, trait=0):
    G_tot = pieces["G_blk"].sum(axis=0)
    Ts, qs = [], []
    for j in range(J + 1):
        jj = 1 if (J == 1 and j == 0) else j                     # base.py:654-655
        G_loo = G_tot - pieces["G_blk"][jj] if jj < J else G_tot
        T, q = normal_equations(plan, ht, pieces["XX"][jj], G_loo, pieces["M"][jj], trait)
        Ts.append(T)
        qs.append(q.ravel())
    return np.array(Ts), np.array(qs)
