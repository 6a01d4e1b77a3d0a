"""GPU: the int8 tcgen05 / TMEM / TMA kernel path (RHE_PATH_TCGEN05) against the golden vectors, the
CUDA-core path and a size-independent property at a larger shape."""
import numpy as np
import pytest

from helpers import load_golden, oracle_problem
from device_model import assemble_all
from test_gpu_parity import make_engine, plan_for, solve_all

pytestmark = pytest.mark.gpu
TC = 1
# absolute floor of the sigma^2 / SE comparison in units of Var(y) for the N ~ 200 goldens (few hundred individuals:
# every component is a small difference of O(N) traces, so fp32 block products move it by ~1e-6; the N-scale cases in
# test_gpu_scale.py use the 1e-7 floor of SURVEY.md §9.2 plus the float64 yardstick)
SIGMA_ATOL = 1e-5
RHE_CASES = ["rhe_cov_binary", "rhe_nocov_mean", "rhe_overlap", "rhe_one_block", "rhe_example_shape",
             "genie_full_cov", "genie_full_nocov", "dom_cov", "dom_nocov"]


@pytest.mark.parametrize("name", RHE_CASES)
def test_tcgen05_T_q_sigma_match_reference(name):
    g = load_golden(name)
    for t in range(g["T"].shape[0]):
        p = oracle_problem(name, trait=t)
        plan = plan_for(p)
        eng, ht, _ = make_engine(p, plan, kernel_path=TC)
        pieces = eng.run()
        eng.close()
        T, q = assemble_all(plan, ht, pieces, p.num_jack)
        np.testing.assert_allclose(T, g["T"][t], rtol=1e-5, atol=1e-6 * np.abs(g["T"][t]).max())
        np.testing.assert_allclose(q, g["q"][t], rtol=1e-5, atol=1e-6 * np.abs(g["q"][t]).max())
        sig = solve_all(T, q)
        vy = float(np.var(p.y))
        ref = np.asarray(g["res_sigma_ests_total"][t])
        J = p.num_jack
        se = np.sqrt((J - 1) * ((sig[:-1] - sig[:-1].mean(0)) ** 2).sum(0) / J)
        from test_gpu_scale import _record, _units
        _record(f"{name}[trait {t}]", {
            "source": "unmodified reference (tools/make_golden.py)", "N": int(p.Z.shape[0]), "M": int(p.annot.shape[0]),
            "J": J, "kernel_path": "tcgen05", "var_y": vy,
            "T_tol_units(rtol 1e-5, atol 1e-6 max|T|)": _units(T, g["T"][t], 1e-5, 1e-6 * np.abs(g["T"][t]).max()),
            "sigma2_total_max_abs/var_y": float(np.max(np.abs(sig[-1] - ref)) / vy),
            "se_max_abs/var_y": float(np.max(np.abs(se - g["res_sig_errs"][t])) / vy)})
        np.testing.assert_allclose(sig[-1], ref, rtol=1e-5, atol=SIGMA_ATOL * vy)
        np.testing.assert_allclose(se, g["res_sig_errs"][t], rtol=1e-4, atol=SIGMA_ATOL * vy)


def test_tcgen05_vectors_match_simt_path():
    """XXz partials and totals of the two kernel paths agree to fixed-point resolution."""
    p = oracle_problem("rhe_cov_binary")
    plan = plan_for(p)
    out = {}
    for path in (0, TC):
        eng, _, _ = make_engine(p, plan, kernel_path=path)
        pieces = eng.run()
        out[path] = (eng.S.cpu().numpy().astype(np.float64), eng.P_all.cpu().numpy().astype(np.float64), pieces)
        eng.close()
    scale = np.abs(out[0][0]).max()
    np.testing.assert_allclose(out[TC][0], out[0][0], rtol=0, atol=2e-6 * scale)
    np.testing.assert_allclose(out[TC][1], out[0][1], rtol=0, atol=2e-6 * scale)
    np.testing.assert_allclose(out[TC][2]["G_blk"], out[0][2]["G_blk"], rtol=1e-6, atol=1e-6 * np.abs(out[0][2]["G_blk"]).max())


def test_tcgen05_larger_shape_against_simt_and_split_invariance():
    """N = 20k, M = 4k, several CTAs per dimension: tensor path == CUDA-core path, and the exact
    integer accumulation makes pass A independent of the jackknife partition (J = 2 vs J = 4 totals)."""
    from pyrhe_b200 import synth
    from pyrhe_b200.assemble import PathPlan
    from pyrhe_b200.engine import RheEngine
    from pyrhe_b200.hostmath import host_terms
    rng = np.random.default_rng(5)
    N, M, K, B = 20_000, 4_096, 4, 10
    packed = synth.pack_counts(synth.random_counts(N, M, rng, missing_rate=0.002))
    annot = synth.random_annot(M, K, rng)
    Z = rng.standard_normal((N, B))
    W = rng.standard_normal((N, 3))
    y = rng.standard_normal((N, 1))
    y -= y.mean()
    plan = PathPlan(model="rhe", K=K, B=B, C=3)
    ht, Y_res = host_terms(plan, Z, W, y, None)
    res = {}
    for path, J, impute in ((0, 4, "binary"), (TC, 4, "binary"), (TC, 4, "mean"), (TC, 2, "mean")):
        eng = RheEngine(plan, n_indv=N, keep=np.ones(N, bool), annot=annot, num_jack=J, impute=impute, seed=3,
                        kernel_path=path)
        eng.set_rhs(Z, W, Y_res)
        eng.load_genotypes(packed)
        pieces = eng.run()
        res[(path, J, impute)] = (pieces, eng.S.cpu().numpy().astype(np.float64))
        eng.close()
    a, b = res[(0, 4, "binary")], res[(TC, 4, "binary")]
    np.testing.assert_allclose(b[0]["XX"], a[0]["XX"], rtol=2e-6)
    np.testing.assert_allclose(b[0]["G_blk"], a[0]["G_blk"], rtol=1e-5, atol=1e-6 * np.abs(a[0]["G_blk"]).max())
    np.testing.assert_allclose(b[1], a[1], rtol=0, atol=2e-6 * np.abs(a[1]).max())
    # totals do not depend on how SNPs are grouped into blocks ("mean" imputation: the binary rule
    # indexes its uniforms by block-local SNP, base.py:510, so it legitimately depends on J)
    Gt4, Gt2 = res[(TC, 4, "mean")][0]["G_blk"].sum(0), res[(TC, 2, "mean")][0]["G_blk"].sum(0)
    np.testing.assert_allclose(Gt2, Gt4, rtol=1e-9, atol=1e-9 * np.abs(Gt4).max())


def test_full_size_block_properties():
    """BASELINE full size in N (500k individuals, two jackknife blocks of 2 000 SNPs generated on the device):
    size-independent properties through the C ABI -- tensor path == CUDA-core path, exact repeatability of the
    integer pass A, and linearity in the right-hand sides (scaling Z by a power of two changes nothing but exponents)."""
    import ctypes as C
    import torch
    from pyrhe_b200 import _lib, synth
    from pyrhe_b200.assemble import PathPlan
    from pyrhe_b200.engine import RheEngine
    from pyrhe_b200.hostmath import host_terms
    rng = np.random.default_rng(11)
    N, M, K, B, J = 500_000, 4_000, 8, 10, 2
    annot = synth.random_annot(M, K, rng)
    Z = rng.standard_normal((N, B))
    W = rng.standard_normal((N, 5))
    y = rng.standard_normal((N, 1))
    y -= y.mean()
    plan = PathPlan(model="rhe", K=K, B=B, C=5)
    ht, Y_res = host_terms(plan, Z, W, y, None)
    lib = _lib.load()

    def run(path, zscale=1.0):
        eng = RheEngine(plan, n_indv=N, keep=np.ones(N, bool), annot=annot, num_jack=J, impute="binary", seed=3,
                        kernel_path=path)
        eng.set_rhs(Z * zscale, W, Y_res)
        eng.alloc_genotypes()
        st = torch.cuda.current_stream()
        for j in eng.own:
            rows, m = eng.block_view(j)
            _lib.check(lib.rhe_synth_genotypes(C.c_void_p(rows.data_ptr()), m, eng.pitch, N, eng.ranges[j][0], 99, 0.001,
                                               C.c_void_p(st.cuda_stream)))
        pieces = eng.run()
        S = eng.S.cpu().numpy()
        eng.close()
        return pieces, S

    simt, S0 = run(0)
    tc1, S1 = run(1)
    tc2, S2 = run(1)
    tc4, S4 = run(1, zscale=4.0)
    np.testing.assert_allclose(tc1["XX"], simt["XX"], rtol=2e-6, atol=1e-8 * np.abs(simt["XX"]).max())
    np.testing.assert_allclose(tc1["G_blk"], simt["G_blk"], rtol=1e-5, atol=1e-6 * np.abs(simt["G_blk"]).max())
    # pass A is exact integer arithmetic; the per-bin Gram then sums fp64 products with atomics, so two runs agree
    # to fp64 round-off (not to the fp32 level a floating-point pass A would give)
    gmax = np.abs(tc1["G_blk"]).max()
    np.testing.assert_allclose(tc1["G_blk"], tc2["G_blk"], rtol=0, atol=1e-12 * gmax)
    # Z -> 4 Z: the Z block of the Gram scales by 16, Z x (W, y) by 4, the rest is unchanged -- exactly
    scale = np.ones(plan.Rs)
    scale[:B] = 4.0
    np.testing.assert_allclose(tc4["G_blk"], tc1["G_blk"] * np.outer(scale, scale), rtol=0, atol=1e-11 * 16 * gmax)
    np.testing.assert_allclose(S4, 4.0 * S1, rtol=0, atol=1e-6 * np.abs(S1).max())


@pytest.mark.parametrize("model", ["rhe_dom", "genie"])
def test_tcgen05_eight_bins_wide_weight_groups(model):
    """K = 8, B = 10 with two weight groups: the accumulators of all bins only fit TMEM with one M-tile per CTA
    (the MT = 1 kernel variant).  Tensor path against the CUDA-core path."""
    from pyrhe_b200 import synth
    from pyrhe_b200.assemble import PathPlan
    from pyrhe_b200.engine import RheEngine
    from pyrhe_b200.hostmath import host_terms
    rng = np.random.default_rng(21)
    N, M, K, B, J = 1_500, 2_400, 8, 10, 3
    packed = synth.pack_counts(synth.random_counts(N, M, rng, missing_rate=0.003))
    annot = synth.random_annot(M, K, rng)
    Z = rng.standard_normal((N, B))
    W = rng.standard_normal((N, 2))
    y = rng.standard_normal((N, 1))
    y -= y.mean()
    env = (rng.random(N) < 0.4).astype(np.float64) if model == "genie" else None
    plan = PathPlan(model=model, K=K, B=B, C=2)
    ht, Y_res = host_terms(plan, Z, W, y, env)
    out = {}
    for path in (0, TC):
        eng = RheEngine(plan, n_indv=N, keep=np.ones(N, bool), annot=annot, num_jack=J, impute="binary", seed=2,
                        kernel_path=path)
        eng.set_rhs(Z, W, Y_res, env)
        eng.load_genotypes(packed)
        pieces = eng.run()
        out[path] = (pieces, eng.S.cpu().numpy().astype(np.float64))
        eng.close()
    a, b = out[0], out[TC]
    # the CUDA-core path sums fp32 products (32-row partials); a few elements of the wide-range dominance
    # weights differ from the exact-integer tensor path at that level
    np.testing.assert_allclose(b[1], a[1], rtol=0, atol=1e-5 * np.abs(a[1]).max())
    np.testing.assert_allclose(b[0]["XX"], a[0]["XX"], rtol=1e-5, atol=1e-7 * np.abs(a[0]["XX"]).max())
    np.testing.assert_allclose(b[0]["G_blk"], a[0]["G_blk"], rtol=1e-5, atol=1e-6 * np.abs(a[0]["G_blk"]).max())


EDGE_SHAPES = [
    # N,   M,   K, B, J, C, missing, dropped individuals, impute, model
    (130,  257, 1, 3, 2, 0, 0.00, (),            "mean",   "rhe"),       # one bin, ragged last block, N % 16 != 0
    (515,  300, 3, 5, 3, 2, 0.20, (0, 7, 514),   "binary", "rhe"),       # heavy missingness, first/last individual dropped
    (1000, 96,  4, 2, 4, 1, 0.01, (),            "binary", "rhe"),       # 24 SNPs per block: far fewer rows than one 128-row stage
    (777,  640, 2, 4, 5, 3, 0.02, (100, 101),    "mean",   "rhe_dom"),   # dominance operand with covariates
    (640,  512, 2, 3, 4, 2, 0.01, (),            "binary", "genie"),     # G + GxE + NxE
    (600, 1280, 20, 10, 4, 2, 0.01, (),          "binary", "rhe"),       # 20 bins: three pass-B launches of <= 8 bins (TMEM columns)
    (520,  960, 11, 9, 3, 0, 0.00, (3,),         "mean",   "rhe_dom"),   # 11 bins x 2 weight groups: bin groups of 4
    (400,  516, 2, 50, 3, 1, 0.01, (),           "binary", "rhe_dom"),   # 50 vectors x 2 weight groups: pass B in column chunks of 32
    (400,  516, 3, 34, 3, 2, 0.01, (5,),         "mean",   "genie"),     # 34 vectors x 2 RHS sets: chunks of 32 + 2
    (300,  384, 8, 50, 3, 0, 0.00, (),           "binary", "rhe"),       # 50 vectors, one weight group: one bin per launch
    (400,  516, 2, 10, 3, 40, 0.01, (),          "binary", "genie"),     # 40 covariates x 2 RHS sets: pass A in two column chunks
]


@pytest.mark.parametrize("path", [0, TC])
@pytest.mark.parametrize("shape", EDGE_SHAPES, ids=[f"N{s[0]}_M{s[1]}_K{s[2]}_{s[9]}" for s in EDGE_SHAPES])
def test_edge_shapes_match_cpu_oracle(shape, path):
    """Ragged / tiny / heavily-missing inputs through both kernel paths against the CPU oracle (T, q within 1e-5)."""
    from oracle import rhe_oracle
    from pyrhe_b200 import synth
    N, M, K, B, J, C, miss, dropped, impute, model = shape
    rng = np.random.default_rng(N * 7 + M)
    counts = synth.random_counts(N, M, rng, missing_rate=miss)
    annot = synth.random_annot(M, K, rng)
    if K > 1:
        annot[: M // J] = 0
        annot[: M // J, 0] = 1                      # the first block has SNPs of bin 0 only: empty bins in a block
    packed = synth.pack_counts(counts)
    n_kept = N - len(dropped)
    Z = rng.standard_normal((n_kept, B))
    W = rng.standard_normal((n_kept, C)) if C else None
    y = rng.standard_normal((n_kept, 1))
    y -= y.mean()
    env = (rng.random(n_kept) < 0.4).astype(np.float64) if model == "genie" else None
    prob = rhe_oracle.OracleProblem(packed=packed, n_indv_original=N, annot=annot, Z=Z, y=y, num_jack=J, W=W,
                                    impute=impute, seed=3, model=model, missing_indv=tuple(dropped), env=env)
    ref = rhe_oracle.run(prob)
    plan = plan_for(prob)
    eng, ht, _ = make_engine(prob, plan, kernel_path=path)
    pieces = eng.run()
    eng.close()
    T, q = assemble_all(plan, ht, pieces, J)
    np.testing.assert_allclose(T, ref["T"], rtol=1e-5, atol=1e-6 * np.abs(ref["T"]).max())
    np.testing.assert_allclose(q, ref["q"], rtol=1e-5, atol=1e-6 * np.abs(ref["q"]).max())


def test_pass_b_two_cta_variant_matches_default(monkeypatch):
    """PYRHE_B200_PASSB_GROUPS=2 (two half-size pass-B CTAs per SM) gives the same pieces as the default layout."""
    p = oracle_problem("rhe_cov_binary")
    plan = plan_for(p)
    eng, _, _ = make_engine(p, plan, kernel_path=TC)
    a = eng.run()
    eng.close()
    monkeypatch.setenv("PYRHE_B200_PASSB_GROUPS", "2")
    eng, _, _ = make_engine(p, plan, kernel_path=TC)
    b = eng.run()
    eng.close()
    np.testing.assert_allclose(a["XX"], b["XX"], rtol=1e-6)
    np.testing.assert_allclose(a["G_blk"], b["G_blk"], rtol=1e-10, atol=1e-9)
