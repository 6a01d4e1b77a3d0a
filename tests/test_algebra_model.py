"""The restructured algorithm (tests/device_model.py) + product host assembly reproduce the
reference's per-jackknife T and q (golden vectors) -- no GPU needed."""
import numpy as np
import pytest

from golden_cases import CASES, SMALL_CASES
from helpers import load_golden, oracle_problem
from device_model import run_model, assemble_all
from pyrhe_b200.assemble import PathPlan
from pyrhe_b200.hostmath import host_terms, binary_fill_values

SMALL = list(SMALL_CASES)


def plan_for(p, Ty=1):
    C = 0 if p.W is None else p.W.shape[1]
    return PathPlan(model=p.model, K=p.annot.shape[1], B=p.Z.shape[1], C=C, Ty=Ty, genie_model=p.genie_model)


@pytest.mark.parametrize("name", SMALL)
def test_model_matches_reference_T_q(name):
    g = load_golden(name)
    for t in range(g["T"].shape[0]):
        p = oracle_problem(name, trait=t)
        plan = plan_for(p)
        ht, Y_res = host_terms(plan, p.Z, p.W, p.y, p.env)
        pieces = run_model(p.packed, p.n_indv_original, p.annot, p.Z, Y_res, p.W, p.env, p.missing_indv,
                           p.num_jack, p.impute, p.seed, plan)
        np.testing.assert_array_equal(pieces["M"], g["M"])
        T, q = assemble_all(plan, ht, pieces, p.num_jack)
        # reference is fp32 on the block products; the model is fp64 -> agreement ~1e-6
        scale = np.abs(g["T"][t]).max()
        np.testing.assert_allclose(T, g["T"][t], rtol=2e-5, atol=2e-6 * scale)
        np.testing.assert_allclose(q, g["q"][t], rtol=2e-5, atol=2e-6 * np.abs(g["q"][t]).max())


def test_model_state_arrays_match_reference():
    """XXz leave-one-out vectors (fp32 in the reference) agree element-wise."""
    for name in ("rhe_cov_binary", "dom_cov", "genie_full_cov"):
        g = load_golden(name)
        t = g["T"].shape[0] - 1
        p = oracle_problem(name, trait=t)
        plan = plan_for(p)
        ht, Y_res = host_terms(plan, p.Z, p.W, p.y, p.env)
        pieces = run_model(p.packed, p.n_indv_original, p.annot, p.Z, Y_res, p.W, p.env, p.missing_indv,
                           p.num_jack, p.impute, p.seed, plan)
        keep = np.ones(p.n_indv_original, bool)
        keep[list(p.missing_indv)] = False
        J = p.num_jack
        L = pieces["S"][None] - pieces["P"]                      # [J, E, B, N0]
        ref = g["XXz"]                                           # [E, J+1, B, N]
        got = np.concatenate([L, pieces["S"][None]], axis=0).transpose(1, 0, 2, 3)[..., keep]
        np.testing.assert_allclose(got, ref, rtol=0, atol=3e-5 * np.abs(ref).max())


def test_binary_fill_rule_matches_reference_decisions():
    g = load_golden("rhe_cov_binary")
    p = oracle_problem("rhe_cov_binary")
    raw, imp = g["geno_raw"], g["geno_imputed"]                 # [N, M], 3 = missing
    from pyrhe_b200.hostmath import block_ranges
    for a, b in block_ranges(raw.shape[1], p.num_jack):
        r = raw[:, a:b]
        fill = binary_fill_values((r == 1).sum(0), (r == 2).sum(0), (r == 3).sum(0), r.shape[0], p.seed)
        has_missing = (r == 3).any(0)
        ref_fill = np.where(has_missing, np.max(np.where(r == 3, imp[:, a:b], 0), axis=0), fill)
        np.testing.assert_array_equal(fill[has_missing], ref_fill[has_missing])


@pytest.mark.parametrize("name", ["rhe_cov_binary", "dom_cov", "genie_full_cov", "genie_full_nocov", "rhe_overlap"])
def test_batched_assembly_equals_scalar_specification(name):
    from pyrhe_b200.assemble import normal_equations_batch
    p = oracle_problem(name)
    plan = plan_for(p)
    ht, Y_res = host_terms(plan, p.Z, p.W, p.y, p.env)
    pieces = run_model(p.packed, p.n_indv_original, p.annot, p.Z, Y_res, p.W, p.env, p.missing_indv,
                       p.num_jack, p.impute, p.seed, plan)
    J = p.num_jack
    if J == 1:
        return
    T, q = assemble_all(plan, ht, pieces, J)
    G_tot = pieces["G_blk"].sum(0)
    G_loo = np.concatenate([G_tot[None] - pieces["G_blk"], G_tot[None]], axis=0)
    Tb, qb = normal_equations_batch(plan, ht, pieces["XX"], G_loo, pieces["M"])
    np.testing.assert_allclose(Tb, T, rtol=1e-12, atol=1e-12 * np.abs(T).max())
    np.testing.assert_allclose(qb, q, rtol=1e-12, atol=1e-12 * np.abs(q).max())
    # the reusable-buffer helper the production tail uses gives the same leave-one-out pieces
    from pyrhe_b200.assemble import loo_grams
    buf = loo_grams(pieces["G_blk"])
    np.testing.assert_allclose(buf, G_loo, rtol=1e-13, atol=1e-13 * np.abs(G_loo).max())
    assert loo_grams(pieces["G_blk"], buf) is buf
