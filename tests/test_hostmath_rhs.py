"""`hostmath.rhs_matrix` (the [Z | W | y_res] (+ env-scaled) operand the engine uploads, base.py:176-178,396-401 /
genie.py:65-67): the one-pass fp32 / padded form `RheEngine.set_rhs` asks for equals the float64 matrix cast and padded
afterwards, bit for bit, with and without dropped individuals."""
import numpy as np
import pytest

from pyrhe_b200.assemble import PathPlan
from pyrhe_b200.hostmath import rhs_matrix


def _plain(plan, Z, W, Y_res, env, keep):
    """The straightforward statement: concatenate, scatter the kept columns, scale the second set by env."""
    n0 = keep.shape[0]
    base = np.concatenate([Z] + ([W] if W is not None else []) + [Y_res], axis=1)
    R = np.zeros((plan.n_sets * plan.Rs, n0))
    rowscale = np.zeros((plan.n_sets, n0))
    R[: plan.Rs, keep] = base.T
    rowscale[0, keep] = 1.0
    if plan.n_sets == 2:
        e = np.asarray(env, dtype=np.float64)
        R[plan.Rs:, keep] = (base * e[:, None]).T
        rowscale[1, keep] = e
    return R, rowscale


@pytest.mark.parametrize("model,n0,drop,n_cov", [("rhe", 20_011, 0, 5), ("rhe", 10_003, 77, 0), ("genie", 9_001, 0, 3),
                                                ("genie", 5_001, 13, 3), ("rhe_dom", 3_000, 5, 2), ("rhe", 7, 0, 1)])
def test_rhs_matrix_one_pass_fp32_equals_cast_of_float64(model, n0, drop, n_cov):
    rng = np.random.default_rng(n0)
    keep = np.ones(n0, dtype=bool)
    keep[rng.choice(n0, drop, replace=False)] = False
    n = int(keep.sum())
    Z = rng.standard_normal((n, 10))
    W = rng.standard_normal((n, n_cov)) if n_cov else None
    Y = rng.standard_normal((n, 2))
    env = (rng.random(n) < 0.4).astype(float) if model == "genie" else None
    plan = PathPlan(model=model, K=3, B=10, C=n_cov, Ty=2)
    R, rowscale = _plain(plan, Z, W, Y, env, keep)
    R64, rs64 = rhs_matrix(plan, Z, W, Y, env, keep)
    assert np.array_equal(R, R64) and np.array_equal(rowscale, rs64)
    width = (n0 + 511) // 512 * 512
    Rp = np.zeros((R.shape[0], width), dtype=np.float32)
    Rp[:, :n0] = R
    rsp = np.zeros((plan.n_sets, width), dtype=np.float32)
    rsp[:, :n0] = rowscale
    R32, rs32 = rhs_matrix(plan, Z, W, Y, env, keep, dtype=np.float32, width=width)
    assert R32.dtype == np.float32 and R32.flags.c_contiguous
    assert np.array_equal(Rp, R32) and np.array_equal(rsp, rs32)
