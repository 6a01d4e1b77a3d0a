"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors of the
unmodified reference and against the CPU oracle.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest
import scipy.linalg

from golden_cases import CASES, SMALL_CASES
from helpers import load_golden, oracle_problem
from device_model import assemble_all
from pyrhe_b200.assemble import PathPlan
from pyrhe_b200.hostmath import host_terms, block_ranges

pytestmark = pytest.mark.gpu
SMALL = list(SMALL_CASES)
PATHS = [0]  # RHE_PATH_SIMT; the tcgen05 path is added in test_gpu_tcgen05.py


def plan_for(p, Ty=1):
    C = 0 if p.W is None else p.W.shape[1]
    return PathPlan(model=p.model, K=p.annot.shape[1], B=p.Z.shape[1], C=C, Ty=Ty, genie_model=p.genie_model)


def make_engine(p, plan, **kw):
    from pyrhe_b200.engine import RheEngine
    keep = np.ones(p.n_indv_original, dtype=bool)
    keep[list(p.missing_indv)] = False
    eng = RheEngine(plan, n_indv=p.n_indv_original, keep=keep, annot=p.annot, num_jack=p.num_jack,
                    impute=p.impute, seed=p.seed, **kw)
    ht, Y_res = host_terms(plan, p.Z, p.W, p.y, p.env)
    eng.set_rhs(p.Z, p.W, Y_res, p.env)
    eng.load_genotypes(p.packed)
    return eng, ht, keep


def solve_all(T, q):
    out = []
    for Tj, qj in zip(T, q):
        Qm, R = scipy.linalg.qr(Tj)
        out.append(np.ravel(scipy.linalg.solve_triangular(R, Qm.T @ qj.reshape(-1, 1))))
    return np.array(out)


@pytest.mark.parametrize("name", ["rhe_cov_binary", "rhe_nocov_mean", "dom_cov"])
def test_decode_and_impute_bit_exact(name):
    """Decoded genotype counts are bit-exact (north_star), raw and after imputation."""
    g = load_golden(name)
    p = oracle_problem(name)
    eng, _, keep = make_engine(p, plan_for(p))
    for j, (a, b) in enumerate(block_ranges(p.annot.shape[0], p.num_jack)):
        raw = eng.decode_block(j, apply_impute=False)            # [m, N0], 3 = missing
        imp = eng.decode_block(j, apply_impute=True)
        np.testing.assert_array_equal(raw[:, keep].T.astype(np.uint8), g["geno_raw"][:, a:b])
        np.testing.assert_array_equal(imp[:, keep].T.astype(np.uint8), g["geno_imputed"][:, a:b])
        st = eng.block_stats(j)
        ref = g["geno_raw"][:, a:b]
        np.testing.assert_array_equal(st, np.stack([(ref == v).sum(0) for v in (0, 1, 2, 3)], axis=1))
    eng.close()


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("name", SMALL)
def test_T_q_sigma_match_reference(name, path):
    """Per-jackknife normal equations, variance components and jackknife SEs: <= 1e-5 relative
    (with the absolute floor SURVEY.md §9.2 motivates) against the unmodified reference."""
    g = load_golden(name)
    for t in range(g["T"].shape[0]):
        p = oracle_problem(name, trait=t)
        plan = plan_for(p)
        eng, ht, _ = make_engine(p, plan, kernel_path=path)
        pieces = eng.run()
        eng.close()
        np.testing.assert_array_equal(pieces["M"], g["M"])
        T, q = assemble_all(plan, ht, pieces, p.num_jack)
        np.testing.assert_allclose(T, g["T"][t], rtol=1e-5, atol=1e-6 * np.abs(g["T"][t]).max())
        np.testing.assert_allclose(q, g["q"][t], rtol=1e-5, atol=1e-6 * np.abs(g["q"][t]).max())
        sig = solve_all(T, q)
        vy = float(np.var(p.y))
        np.testing.assert_allclose(sig[-1], g["res_sigma_ests_total"][t], rtol=1e-5, atol=1e-5 * vy)
        J = p.num_jack
        se = np.sqrt((J - 1) * ((sig[:-1] - sig[:-1].mean(0)) ** 2).sum(0) / J)
        np.testing.assert_allclose(se, g["res_sig_errs"][t], rtol=1e-4, atol=1e-5 * vy)


@pytest.mark.parametrize("name", ["rhe_cov_binary", "dom_cov", "genie_full_cov"])
def test_state_vectors_match_reference(name):
    """The leave-one-out XXz vectors (S - P_j) against the reference's fp32 state arrays."""
    g = load_golden(name)
    t = g["T"].shape[0] - 1
    p = oracle_problem(name, trait=t)
    plan = plan_for(p)
    eng, _, keep = make_engine(p, plan)
    eng.run()
    S = eng.S.cpu().numpy()[..., : p.n_indv_original][..., keep].astype(np.float64)
    P = eng.P_all.cpu().numpy()[..., : p.n_indv_original][..., keep].astype(np.float64)
    eng.close()
    got = np.concatenate([S[None] - P, S[None]], axis=0).transpose(1, 0, 2, 3)
    ref = g["XXz"]
    np.testing.assert_allclose(got, ref, rtol=0, atol=3e-5 * np.abs(ref).max())


def test_streaming_policy_equals_stored_partials():
    p = oracle_problem("rhe_cov_binary")
    plan = plan_for(p)
    eng, _, _ = make_engine(p, plan)
    a = eng.run()
    eng.close()
    eng, _, _ = make_engine(p, plan, store_partials=False)
    b = eng.run()
    eng.close()
    np.testing.assert_allclose(a["XX"], b["XX"], rtol=1e-12)
    np.testing.assert_allclose(a["G_blk"], b["G_blk"], rtol=1e-10, atol=1e-9)


def test_matches_cpu_oracle_on_fresh_seeded_inputs():
    """Same seeded inputs through the CPU oracle and the CUDA path (no golden file involved)."""
    from oracle import rhe_oracle
    from pyrhe_b200 import synth
    rng = np.random.default_rng(123)
    N, M, K, B, J = 333, 640, 2, 4, 5
    counts = synth.random_counts(N, M, rng, missing_rate=0.01)
    packed = synth.pack_counts(counts)
    annot = synth.random_annot(M, K, rng)
    Z = rng.standard_normal((N, B))
    W = rng.standard_normal((N, 2))
    y = rng.standard_normal((N, 1))
    y -= y.mean()
    prob = rhe_oracle.OracleProblem(packed=packed, n_indv_original=N, annot=annot, Z=Z, y=y, num_jack=J, W=W,
                                    impute="binary", seed=5, model="rhe")
    ref = rhe_oracle.run(prob)
    plan = plan_for(prob)
    eng, ht, _ = make_engine(prob, plan)
    pieces = eng.run()
    eng.close()
    T, q = assemble_all(plan, ht, pieces, J)
    np.testing.assert_allclose(T, ref["T"], rtol=1e-5, atol=1e-6 * np.abs(ref["T"]).max())
    np.testing.assert_allclose(q, ref["q"], rtol=1e-5, atol=1e-6 * np.abs(ref["q"]).max())
