"""GPU: the reference's block interface for extenders, the reference-named state arrays, the bounded-memory ingest
ring, and the edge cases the reference's own code handles (or crashes on)."""
import os

import numpy as np
import pytest

from helpers import case_dataset, load_golden, oracle_problem
from test_api_cpu import build_model
from test_gpu_parity import make_engine, plan_for, solve_all
from device_model import assemble_all

pytestmark = pytest.mark.gpu


def _capture_T_q(model):
    Ts, qs = [], []
    orig = model.setup_lhs_rhs_jackknife

    def capture(j, trace_sums, is_streaming=False):
        T, q = orig(j, trace_sums, is_streaming)
        Ts.append(np.array(T))
        qs.append(np.array(q).ravel())
        return T, q

    model.setup_lhs_rhs_jackknife = capture
    return Ts, qs


def _extender_class(streaming):
    """A model written the way the reference's own RHE / StreamingRHE are written (rhe.py:13-22,
    streaming_rhe.py:13-43): per-bin standardisation and the four `_compute_*` helpers filling the state arrays."""
    import pyrhe.models as models

    if not streaming:
        class MyRHE(models.RHE):
            def pre_compute_jackknife_bin(self, j, all_gen):
                for k, X_kj in enumerate(all_gen):
                    X_kj = self.standardize_geno(X_kj)
                    self.M[j][k] = self.M[self.num_jack][k] - X_kj.shape[1]
                    for b in range(self.num_random_vec):
                        self.XXz[k, j, b, :] = self._compute_XXz(b, X_kj)
                        if self.use_cov:
                            self.UXXz[k, j, b, :] = self._compute_UXXz(self.XXz[k][j][b])
                            self.XXUz[k, j, b, :] = self._compute_XXUz(b, X_kj)
                    self.yXXy[k][j] = self._compute_yXXy(X_kj, y=self.pheno)[0][0]
        return MyRHE

    class MyStreamingRHE(models.StreamingRHE):
        def pre_compute_jackknife_bin(self, j, all_gen, worker_num):
            for k, X_kj in enumerate(all_gen):
                X_kj = self.standardize_geno(X_kj)
                self.M[j][k] = self.M[self.num_jack][k] - X_kj.shape[1]
                for b in range(self.num_random_vec):
                    xxz = self._compute_XXz(b, X_kj)
                    self.XXz[k][worker_num][b] += xxz
                    if self.use_cov:
                        # the intended per-block term (the reference adds U @ running-sum here)
                        self.UXXz[k][worker_num][b] += self._compute_UXXz(xxz)
                        self.XXUz[k][worker_num][b] += self._compute_XXUz(b, X_kj)
                self.yXXy[k][worker_num] += self._compute_yXXy(X_kj, y=self.pheno)[0][0]

        def pre_compute_jackknife_bin_pass_2(self, j, all_gen):
            last = j == self.num_jack
            for k in range(self.num_estimates):
                X_kj = self.standardize_geno(all_gen[k]) if not last else 0
                for b in range(self.num_random_vec):
                    XXz_kb = self._compute_XXz(b, X_kj) if not last else 0
                    if self.use_cov:
                        self.UXXz[k][1][b] = self.UXXz[k][0][b] - (self._compute_UXXz(XXz_kb) if not last else 0)
                        self.XXUz[k][1][b] = self.XXUz[k][0][b] - (self._compute_XXUz(b, X_kj) if not last else 0)
                    self.XXz[k][1][b] = self.XXz[k][0][b] - XXz_kb
                yk = self._compute_yXXy(X_kj, y=self.pheno)[0][0] if not last else 0
                self.yXXy[k][1] = self.yXXy[k][0] - yk
    return MyStreamingRHE


@pytest.mark.parametrize("streaming", [False, True])
@pytest.mark.parametrize("name", ["rhe_cov_binary", "rhe_nocov_mean"])
def test_extender_block_hook_is_driven_like_the_reference(name, streaming):
    """A subclass overriding `pre_compute_jackknife_bin` (and `_pass_2`) is no longer ignored: Base.pre_compute
    detects it and runs the reference's per-block loop (base.py:503-527, base_streaming.py:85-144) over GPU-decoded
    blocks; T, q, sigma and the state arrays reproduce the goldens of the unmodified reference."""
    from pyrhe.src.util import Logger
    g = load_golden(name)
    case, paths = case_dataset(name)
    cls = _extender_class(streaming)
    kw = dict(case["kwargs"])
    kw.update(paths)
    model = cls(model="rhe", log=Logger(suppress=True, debug_mode=False), multiprocessing=False, device="cuda",
                num_workers=1, **kw)
    for t in range(model.num_traits):
        Ts, qs = _capture_T_q(model)
        res = model(trait=t)
        assert model._hook_mode
        T, q = np.array(Ts), np.array(qs)
        np.testing.assert_array_equal(model.M, g["M"])
        np.testing.assert_allclose(T, g["T"][t], rtol=1e-5, atol=1e-6 * np.abs(g["T"][t]).max())
        np.testing.assert_allclose(q, g["q"][t], rtol=1e-5, atol=1e-6 * np.abs(g["q"][t]).max())
        np.testing.assert_allclose(res["sigma_ests_total"], g["res_sigma_ests_total"][t], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(res["sig_errs"], g["res_sig_errs"][t], rtol=1e-4, atol=2e-6)
        if not streaming and t == model.num_traits - 1:
            for key in ("XXz", "UXXz", "XXUz", "yXXy"):
                if key in g.files:
                    ref = g[key]
                    np.testing.assert_allclose(getattr(model, key), ref, rtol=0, atol=3e-5 * np.abs(ref).max(),
                                               err_msg=key)


@pytest.mark.parametrize("name,streaming", [("rhe_cov_binary", False), ("rhe_cov_binary", True), ("dom_cov", False),
                                            ("genie_full_cov", False), ("rhe_nocov_mean", False)])
def test_state_arrays_of_builtin_models(name, streaming):
    """`model.XXz / UXXz / XXUz / yXXy` of the fused built-in models, materialised on demand in the reference's shapes
    and post-`aggregate` meaning (base.py:419-429; docs/api/base.rst:75-116), against the reference's own arrays."""
    g = load_golden(name)
    model, _ = build_model(name, streaming=streaming)
    last = model.num_traits - 1
    for t in range(model.num_traits):
        model._finalize = lambda: None               # keep the object alive for inspection after the last trait
        model(trait=t)
    assert not model._hook_mode
    E, J, B, N = model.num_estimates, model.num_jack, model.num_random_vec, model.num_indv
    assert model.XXz.shape == (E, J + 1, B, N) and model.yXXy.shape == (E, J + 1)
    for key in ("XXz", "yXXy", "UXXz", "XXUz"):
        if key not in g.files:
            with pytest.raises(AttributeError):
                getattr(model, key)
            continue
        ref = g[key]
        np.testing.assert_allclose(getattr(model, key), ref, rtol=0, atol=3e-5 * np.abs(ref).max(), err_msg=key)
    assert last >= 0


def test_ingest_ring_bounds_memory_and_matches_resident(tmp_path):
    """A `.bed` share streamed through a ring of TWO block slots (stored partials, and the streaming policy that
    streams the file a second time) gives exactly the pieces of the all-resident run: HBM use is bounded by the ring,
    not by the size of the file (base.py:338-345 reads one block at a time)."""
    p = oracle_problem("rhe_cov_binary")
    plan = plan_for(p)
    eng, _, _ = make_engine(p, plan)
    ref = eng.run()
    resident_bytes = eng.genotype_bytes()
    eng.close()
    bed_path = os.path.join(tmp_path, "x.bed")
    with open(bed_path, "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]))
        f.write(np.ascontiguousarray(p.packed).tobytes())
    mm = np.memmap(bed_path, dtype=np.uint8, mode="r", offset=3, shape=p.packed.shape)
    for store, source in ((True, mm), (False, mm), (True, np.array(p.packed))):
        from pyrhe_b200.engine import RheEngine
        from pyrhe_b200.hostmath import host_terms
        keep = np.ones(p.n_indv_original, dtype=bool)
        keep[list(p.missing_indv)] = False
        eng = RheEngine(plan, n_indv=p.n_indv_original, keep=keep, annot=p.annot, num_jack=p.num_jack,
                        impute=p.impute, seed=p.seed, store_partials=store)
        _, Y_res = host_terms(plan, p.Z, p.W, p.y, p.env)
        eng.set_rhs(p.Z, p.W, Y_res, p.env)
        st = eng.stream_genotypes(source, n_workers=3, ring_host=2, ring_blocks=2)
        out = eng.run(upload=st)
        assert eng.ring_blocks == 2 and eng.genotype_bytes() < resident_bytes
        assert st.passes == (1 if store else 2)
        assert st.bytes_staged == st.passes * p.packed.size
        st.close()
        eng.close()
        np.testing.assert_allclose(out["XX"], ref["XX"], rtol=1e-12)
        np.testing.assert_allclose(out["G_blk"], ref["G_blk"], rtol=1e-10, atol=1e-9)


def test_resident_counts_equal_recounting():
    """Allele counts taken once at ingest (`rhe_block_stats`) and handed to every `rhe_block_accumulate` give the same
    pieces as re-counting inside the call (the three-read path of round 1) -- to fp64 round-off: the integer pass A
    is bit-identical, the Gram kernels then sum fp64 products with atomics in arbitrary order."""
    p = oracle_problem("dom_cov")
    plan = plan_for(p)
    outs = []
    for resident in (True, False):
        eng, _, _ = make_engine(p, plan)
        eng.use_resident_counts = resident
        eng.use_fast_layout = False                    # same pass-B kernel in both runs (the copies need the counts)
        outs.append(eng.run())
        eng.close()
    np.testing.assert_allclose(outs[0]["XX"], outs[1]["XX"], rtol=1e-12)
    np.testing.assert_allclose(outs[0]["G_blk"], outs[1]["G_blk"], rtol=1e-12, atol=1e-12 * np.abs(outs[1]["G_blk"]).max())


@pytest.mark.parametrize("path", [0, 1])
def test_unannotated_block_followed_by_annotated_block(path):
    """Region-restricted annotation: the SNPs of the first TWO jackknife blocks sit in no bin at all.  (Round 1 keyed
    its per-block metadata cache on a pointer that such empty blocks share with their successor; plans are now
    explicit handles.)  Against the CPU oracle."""
    from oracle import rhe_oracle
    from pyrhe_b200 import synth
    rng = np.random.default_rng(77)
    N, M, K, B, J = 300, 600, 3, 4, 6
    packed = synth.pack_counts(synth.random_counts(N, M, rng, missing_rate=0.01))
    annot = synth.random_annot(M, K, rng)
    annot[: 2 * (M // J)] = 0
    Z = rng.standard_normal((N, B))
    W = rng.standard_normal((N, 2))
    y = rng.standard_normal((N, 1))
    y -= y.mean()
    prob = rhe_oracle.OracleProblem(packed=packed, n_indv_original=N, annot=annot, Z=Z, y=y, num_jack=J, W=W,
                                    impute="binary", seed=4, model="rhe")
    ref = rhe_oracle.run(prob)
    plan = plan_for(prob)
    eng, ht, _ = make_engine(prob, plan, kernel_path=path)
    pieces = eng.run()
    eng.close()
    T, q = assemble_all(plan, ht, pieces, J)
    np.testing.assert_allclose(T, ref["T"], rtol=1e-5, atol=1e-6 * np.abs(ref["T"]).max())
    np.testing.assert_allclose(q, ref["q"], rtol=1e-5, atol=1e-6 * np.abs(ref["q"]).max())


def test_monomorphic_snp_fails_like_the_reference(tmp_path):
    """A SNP with no variation: `standardize_geno` divides by sqrt(mu (1 - mu / 2)) = 0 (base.py:291-296), the
    reference's statistics become NaN and its QR solve raises `ValueError` (scipy checks finiteness).  The CUDA path
    propagates the NaN the same way: the model call fails with the same exception instead of returning numbers."""
    import pyrhe.models as models
    from pyrhe.src.util import Logger
    from pyrhe_b200.synth import make_dataset
    paths = make_dataset(str(tmp_path), "mono", N=300, M=240, K=2, seed=5, n_cov=2, monomorphic=(17,))
    model = models.RHE(model="rhe", log=Logger(suppress=True, debug_mode=False), multiprocessing=False, device="cuda",
                       num_workers=1, num_jack=4, num_random_vec=4, geno_impute_method="mean", seed=2, **paths)
    with pytest.raises(ValueError):
        model(trait=0)
    pc = model._pieces
    assert np.isnan(pc["G_blk"]).any() and np.isnan(pc["XX"]).any()


def test_covariate_mean_imputation_keeps_the_individuals(tmp_path):
    """`cov_impute_method="mean"` with NA cells (the reference cannot run this file, see tools/golden_cases.py): the
    cells take their column mean and the individuals stay.  Equals a run on a file where the means were written in."""
    import pyrhe.models as models
    from pyrhe.src.util import Logger
    from pyrhe_b200.synth import make_dataset
    cells = ((3, 1), (100, 2), (239, 0))
    a = make_dataset(os.path.join(tmp_path, "a"), "c", N=240, M=360, K=2, seed=62, n_cov=3, cov_missing=cells)
    b = make_dataset(os.path.join(tmp_path, "b"), "c", N=240, M=360, K=2, seed=62, n_cov=3)
    import pandas as pd
    df = pd.read_csv(b["cov_file"], sep=r"\s+")
    ref = df.copy()
    for i, c in cells:
        col = f"cov{c}"
        ref.loc[i, col] = df[col].drop(index=[r for r, cc in cells if cc == c]).mean()
    ref.to_csv(b["cov_file"], sep=" ", index=False, float_format="%.17g")
    kw = dict(num_jack=5, num_random_vec=4, geno_impute_method="binary", seed=10)
    out = []
    for paths, method in ((a, "mean"), (b, "ignore")):
        m = models.RHE(model="rhe", log=Logger(suppress=True, debug_mode=False), multiprocessing=False, device="cuda",
                       num_workers=1, cov_impute_method=method, **kw, **paths)
        assert m.num_indv == 240
        out.append(m(trait=0))
    for key in out[0]:
        np.testing.assert_allclose(out[0][key], out[1][key], rtol=1e-9, atol=1e-12, err_msg=key)


@pytest.mark.parametrize("name", ["rhe_cov_binary", "rhe_nocov_mean", "rhe_overlap", "rhe_one_block", "dom_cov", "dom_nocov",
                                  "genie_full_cov", "genie_full_nocov", "rhe_example_shape"])
def test_individual_major_fast_path_equals_gather(name):
    """Pass B fed from tensor memory off the block's individual-major copy (`rhe_block_transpose`, k_tc_pass_b2) against
    the gather kernel on the SNP-major rows: the same exact integer accumulation and the same fp64 epilogue (RHE-DOM's
    two operands accumulate into one tensor-memory accumulator on both paths), so the block partials agree to the last
    fp32 bit wherever the per-bin mean term does -- that term is an fp64 atomic sum, whose last bit may differ from run
    to run and move a partial that sits on a rounding boundary by one fp32 ulp."""
    p = oracle_problem(name)
    plan = plan_for(p)
    out = {}
    for fast in (True, False):
        eng, _, _ = make_engine(p, plan, kernel_path=1)
        assert len(eng.gt) == len(eng.own), "every block of these small cases gets its individual-major copy"
        eng.use_fast_layout = fast
        pieces = eng.run()
        out[fast] = (pieces, eng.P_all.cpu().numpy(), eng.S.cpu().numpy())
        eng.close()
    a, b = out[True], out[False]
    differ = a[1] != b[1]
    assert differ.mean() < 1e-5
    np.testing.assert_allclose(a[1], b[1], rtol=2.4e-7, atol=0)
    np.testing.assert_allclose(a[2], b[2], rtol=0, atol=2e-6 * np.abs(b[2]).max())
    np.testing.assert_allclose(a[0]["XX"], b[0]["XX"], rtol=1e-6)
    np.testing.assert_array_equal(a[0]["G_blk"], a[0]["G_blk"])


@pytest.mark.parametrize("n_est", [1, 7, 8, 9, 16, 17, 23, 24])
def test_leave_one_out_gram_against_numpy(n_est):
    """`rhe_loo_gram_multi` (base.py:578-581 after the aggregate of base.py:483-486): out[j][a][c] = <S_a - P_ja, S_c - P_jc>
    for several stored blocks per launch, every row tiling of the FP64 tensor-core kernel (one to three 8-row tiles, with and
    without the trailing row that GENIE's 2 K + 1 estimates add) against float64 numpy."""
    import ctypes as C
    import torch
    from pyrhe_b200 import _lib
    p = oracle_problem("rhe_nocov_mean")
    eng, _, _ = make_engine(p, plan_for(p), kernel_path=1)
    lib = _lib.load()
    rng = np.random.default_rng(n_est)
    length, n_blocks = 16 * 37, 7
    S = rng.standard_normal((n_est, length)).astype(np.float32) * 3
    P = rng.standard_normal((n_blocks, n_est, length)).astype(np.float32)
    dS, dP = torch.from_numpy(S).cuda(), torch.from_numpy(P).cuda()
    out = torch.full((n_blocks, n_est, n_est), 123.0, dtype=torch.float64, device="cuda")
    _lib.check(lib.rhe_loo_gram_multi(eng._ctx, _lib.ptr(dS), _lib.ptr(dP), n_est * length, n_blocks, n_est, length,
                                      _lib.ptr(out), n_est * n_est, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    D = S[None].astype(np.float64) - P.astype(np.float64)
    want = np.einsum("jal,jcl->jac", D, D)
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-12, atol=1e-9)
    eng.close()


@pytest.mark.parametrize("name", ["rhe_cov_binary", "rhe_nocov_mean", "rhe_overlap", "dom_cov", "genie_full_cov", "rhe_example_shape"])
def test_retiled_rows_give_the_same_block_results(name):
    """`retile=True`: once the counts and the individual-major copy of a block are taken, its SNP-major rows are
    rewritten in place as contiguous 16 KB boxes of imputed counts (`rhe_block_retile`) and pass A reads those
    (`k_tc_pass_a<., 1>`: mask-and-shift decode, no per-SNP table).  Pass A is exact integer arithmetic on the same
    values in the same order, so the per-bin Grams agree to fp64 round-off of their atomic sums; the decode / recount /
    gather-kernel hooks of such an engine refuse to run instead of misreading the rows."""
    from pyrhe_b200 import _lib
    p = oracle_problem(name)
    plan = plan_for(p)
    out = {}
    for retile in (False, True):
        eng, _, _ = make_engine(p, plan, kernel_path=1, retile=retile)
        assert len(eng._tiled) == (len(eng.own) if retile else 0)
        pieces = eng.run()
        out[retile] = (pieces, eng.S.cpu().numpy())
        if retile:
            with pytest.raises(_lib.RheError):
                eng.decode_block(eng.own[0], apply_impute=False)
            eng.use_fast_layout = False
            with pytest.raises(_lib.RheError):
                eng.run()
        eng.close()
    a, b = out[True], out[False]
    gmax = np.abs(b[0]["G_blk"]).max()
    np.testing.assert_allclose(a[0]["G_blk"], b[0]["G_blk"], rtol=0, atol=1e-12 * gmax)
    np.testing.assert_allclose(a[0]["XX"], b[0]["XX"], rtol=1e-6)
    np.testing.assert_allclose(a[1], b[1], rtol=0, atol=2e-6 * np.abs(b[1]).max())


def test_totals_summed_from_stored_partials_are_reproducible():
    """`sum_stored_partials`: S = sum_j P_j in block order (`rhe_sum_partials`) instead of RED from every pass B -- the
    same totals to fp32 round-off, and identical from run to run."""
    p = oracle_problem("rhe_cov_binary")
    plan = plan_for(p)
    eng, _, _ = make_engine(p, plan, kernel_path=1)
    ref = eng.run()
    S_red = eng.S.cpu().numpy().copy()
    eng.sum_stored_partials = True
    a = eng.run()
    S1 = eng.S.cpu().numpy().copy()
    P1 = eng.P_all.cpu().numpy()
    b = eng.run()
    S2 = eng.S.cpu().numpy()
    eng.close()
    np.testing.assert_allclose(S1, S_red, rtol=0, atol=2e-6 * np.abs(S_red).max())
    np.testing.assert_allclose(S1, P1.astype(np.float64).sum(axis=0), rtol=0, atol=2e-6 * np.abs(S_red).max())
    np.testing.assert_array_equal(S1, S2)
    np.testing.assert_allclose(a["XX"], b["XX"], rtol=1e-12)   # (the Gram's own fp64 atomics are not ordered)
    np.testing.assert_allclose(a["XX"], ref["XX"], rtol=2e-6)


def test_retile_is_refused_beyond_two_million_individuals():
    """Pass A on re-tiled rows feeds 4 x the count to the MMA, so its int32 accumulation is exact only for
    8 x 128 x N < 2^31: `rhe_block_tiled_bytes` answers 0 from 2^21 individuals on and the engine keeps the PLINK rows."""
    from pyrhe_b200.assemble import PathPlan
    from pyrhe_b200.engine import RheEngine
    from pyrhe_b200 import synth
    rng = np.random.default_rng(5)
    N, M, K, B, J = (1 << 21) + 7, 24, 1, 2, 2
    packed = synth.pack_counts(synth.random_counts(N, M, rng))
    eng = RheEngine(PathPlan(model="rhe", K=K, B=B, C=0), n_indv=N, keep=np.ones(N, bool), annot=np.ones((M, K), dtype=np.int64),
                    num_jack=J, impute="mean", seed=0, kernel_path=1, retile=True)
    assert int(eng.lib.rhe_block_tiled_bytes(eng._ctx, eng._plans[0])) == 0
    Z = rng.standard_normal((N, B))
    y = rng.standard_normal((N, 1))
    eng.set_rhs(Z, None, y - y.mean())
    eng.load_genotypes(packed)
    assert len(eng.gt) == J and not eng._tiled and eng._retile_scratch is None
    out = eng.run()
    eng.close()
    # <X X^T z, z> summed over the vectors against float64 numpy on the decoded counts
    g = np.where(np.unpackbits(packed[:, :, None], axis=2, bitorder="little").reshape(M, -1, 2)[:, :N].dot([1, 2]) == 2, 1,
                 np.where(np.unpackbits(packed[:, :, None], axis=2, bitorder="little").reshape(M, -1, 2)[:, :N].dot([1, 2]) == 3, 2, 0))
    mu = g.mean(axis=1, keepdims=True)
    X = ((g - mu) / np.sqrt(mu * (1 - mu / 2))).T
    XXz = X @ (X.T @ Z)
    np.testing.assert_allclose(out["XX"][J, 0, 0], np.sum(XXz * XXz), rtol=1e-5)
