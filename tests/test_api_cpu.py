"""CPU-side checks: the drop-in API surface, host logic and the C-ABI library (no GPU compute)."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

from golden_cases import CASES
from helpers import ROOT, case_dataset, load_golden


def build_model(name, **extra):
    import pyrhe.models as models
    from pyrhe.src.util import Logger
    case, paths = case_dataset(name)
    cls = getattr(models, ("Streaming" if extra.pop("streaming", False) else "") + case["model"])
    log = Logger(suppress=True, debug_mode=False)
    kw = dict(case["kwargs"])
    kw.update(paths)
    kw.update(extra)
    model_name = {"RHE": "rhe", "RHE_DOM": "rhe_dom", "GENIE": "genie"}[case["model"]]
    return cls(model=model_name, log=log, multiprocessing=False, device="cpu", num_workers=1, **kw), log


def test_import_paths_and_mro():
    import pyrhe.models as m
    from pyrhe.src.base import Base, StreamingBase
    from pyrhe.src.models.genie import GENIE, StreamingGENIE
    from pyrhe.src.models.rhe import RHE, StreamingRHE
    from pyrhe.src.models.rhe_dom import RHE_DOM, StreamingRHE_DOM
    assert m.RHE is RHE and m.StreamingGENIE is StreamingGENIE
    assert StreamingRHE.__mro__[1:3] == (RHE, StreamingBase)
    assert StreamingRHE_DOM.__mro__[1:3] == (RHE_DOM, StreamingBase)
    assert StreamingGENIE.__mro__[1:3] == (GENIE, StreamingBase)
    assert issubclass(StreamingBase, Base)
    params = inspect.signature(Base.__init__).parameters
    for kw in ("model", "geno_file", "annot_file", "pheno_file", "cov_file", "num_bin", "num_jack", "num_random_vec",
               "geno_impute_method", "cov_impute_method", "cov_one_hot_conversion", "categorical_threshhold",
               "device", "cuda_num", "num_workers", "multiprocessing", "seed", "get_trace", "trace_dir",
               "samp_prev", "pop_prev", "log"):
        assert kw in params
    for hook in ("get_num_estimates", "get_M_last_row", "pre_compute_jackknife_bin", "b_trace_calculation", "run",
                 "pre_compute", "aggregate", "setup_lhs_rhs_jackknife", "estimate", "estimate_error",
                 "compute_h2_nonoverlapping", "compute_h2_overlapping", "compute_enrichment", "get_trace_summary",
                 "solve_linear_equation", "solve_linear_qr", "regress_pheno", "_distribute_work"):
        assert hasattr(Base, hook)
    assert hasattr(StreamingBase, "pre_compute_jackknife_bin_pass_2")


@pytest.mark.parametrize("name", ["rhe_cov_binary", "dom_nocov", "genie_full_cov"])
def test_constructor_reproduces_reference_state(name):
    """Same files + integer seed -> same Z (draw order after np.random.seed), same filtering, same
    header log lines as the unmodified reference."""
    g = load_golden(name)
    model, log = build_model(name)
    np.testing.assert_array_equal(model.all_zb, g["Z"])
    np.testing.assert_array_equal(np.asarray(model.missing_indv), g["missing_indv"])
    assert model.num_indv == int(g["num_indv"])
    np.testing.assert_array_equal(np.asarray(model.get_M_last_row()), g["M"][-1])
    ref_header = str(g["log"]).split("*****\nOUTPUT FOR TRAIT")[0]
    assert "".join(log.msgs) == ref_header


def test_seed_string_is_coerced_and_annotation_generated(tmp_path, monkeypatch):
    """CLI seeds arrive as str (Q6); `annot_file=None` draws the annotation from the global RNG
    BEFORE Z (base.py:112,176), so Z must equal what the reference's order produces."""
    from pyrhe.models import RHE
    from pyrhe.src.util import Logger
    monkeypatch.chdir(tmp_path)
    _, paths = case_dataset("rhe_nocov_mean")
    m = RHE(model="rhe", geno_file=paths["geno_file"], pheno_file=paths["pheno_file"], annot_file=None, num_bin=3,
            num_jack=4, num_random_vec=2, seed="5", log=Logger(suppress=True, debug_mode=False), num_workers=1)
    np.random.seed(5)
    expect_bins = np.random.randint(0, 3, size=m.num_snp)
    expect_Z = np.random.randn(m.num_indv, 2)
    np.testing.assert_array_equal(np.argmax(m.annot_matrix, axis=1), expect_bins)
    np.testing.assert_array_equal(m.all_zb, expect_Z)
    with pytest.raises(ValueError):
        RHE(model="rhe", geno_file=paths["geno_file"], pheno_file=paths["pheno_file"], annot_file=None, num_bin=None,
            num_jack=4, log=Logger(suppress=True, debug_mode=False), num_workers=1)


def test_c_abi_library_exports_every_declared_symbol():
    from pyrhe_b200 import _lib
    header = open(os.path.join(ROOT, "include", "pyrhe_b200.h")).read()
    declared = set(re.findall(r"\b(rhe_[a-z0-9_]+)\s*\(", header)) - {"rhe_ctx"}
    lib = _lib.load()
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.rhe_version() == 4
    # argument validation happens before any CUDA call
    cfg = _lib.RheConfig(device=0, n_indv=10, n_kept=10, pitch_bytes=100, n_cols_set=4, n_sets=1, n_ops=1, n_vec=2,
                         n_bins=1, max_block_snps=8, impute_binary=0, kernel_path=0)
    ctx = ctypes.c_void_p()
    assert lib.rhe_ctx_create(ctypes.byref(ctx), ctypes.byref(cfg)) == -1
    assert b"pitch_bytes" in lib.rhe_last_error()


def test_tcgen05_capability_is_answered_by_the_library():
    """The layout limits of the tensor kernels live in ONE place (rhe_tc_supported, also used by rhe_ctx_create): the
    Python side asks instead of mirroring them.  Pure host arithmetic -- no device is touched."""
    from pyrhe_b200 import _lib
    from pyrhe_b200.assemble import PathPlan
    assert _lib.tcgen05_supported(PathPlan(model="rhe", K=8, B=10, C=5))
    assert _lib.tcgen05_supported(PathPlan(model="rhe_dom", K=8, B=10, C=5))
    assert _lib.tcgen05_supported(PathPlan(model="genie", K=8, B=10, C=5))
    assert _lib.tcgen05_supported(PathPlan(model="rhe", K=20, B=10, C=2))         # bin groups
    # shapes beyond one launch's TMEM / shared-memory layout run in column chunks instead of leaving the tensor path:
    # 64 vectors, the reference's 50 vectors with two weight groups, 40 covariates with two RHS sets
    assert _lib.tcgen05_supported(PathPlan(model="rhe", K=2, B=64, C=0))
    assert _lib.tcgen05_supported(PathPlan(model="rhe_dom", K=8, B=50, C=5))
    assert _lib.tcgen05_supported(PathPlan(model="genie", K=8, B=50, C=40))
    assert _lib.tcgen05_unsupported_reason(PathPlan(model="genie", K=8, B=50, C=40)) == ""
    bad = _lib.plan_config(PathPlan(model="rhe", K=1, B=2, C=0), pitch_bytes=100)   # invalid config -> 0, not a crash
    assert _lib.load().rhe_tc_supported(ctypes.byref(bad)) == 0


def test_product_has_no_cpu_fallback_and_never_imports_the_oracle():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    model, _ = build_model("rhe_nocov_mean")
    with pytest.raises(Exception) as ei:
        model(trait=0)
    assert "CUDA" in str(ei.value)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pyrhe_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", "").lower() or f == "synth.py", f


def test_host_statistics_match_reference_given_reference_T_q():
    """Estimator tail (solve, SE, h2, enrichment) from the golden T, q reproduces the golden result dict."""
    from pyrhe_b200 import stats
    g = load_golden("rhe_overlap")
    T, q = g["T"][0], g["q"][0]
    J, E = T.shape[0] - 1, T.shape[1] - 1
    sig = np.array([stats.solve(T[j], q[j].reshape(-1, 1), "QR") for j in range(J + 1)])
    np.testing.assert_allclose(sig[-1], g["res_sigma_ests_total"][0], rtol=1e-12)
    np.testing.assert_allclose(stats.jackknife_se(sig[:-1], J), g["res_sig_errs"][0], rtol=1e-9)
    h2 = stats.h2_nonoverlapping(sig, E)
    np.testing.assert_allclose(h2[-1], g["res_h2_total"][0], rtol=1e-12)
    en = stats.enrichment(h2, g["M"], E)
    np.testing.assert_allclose(en[-1], g["res_enrichment_total"][0], rtol=1e-12)
    model, _ = build_model("rhe_overlap")
    from pyrhe_b200.hostmath import block_ranges
    tot, blk = stats.bin_cooccurrence(model.annot_matrix, block_ranges(model.num_snp, J))
    h2o = stats.h2_overlapping(sig, g["M"], tot, blk, E)
    np.testing.assert_allclose(h2o[-1], g["res_h2_total_overlap"][0], rtol=1e-10)
    np.testing.assert_allclose(stats.jackknife_se(h2o[:-1], J), g["res_h2_errs_overlap"][0], rtol=1e-8, atol=1e-14)


def test_block_partition_matches_reference_rule():
    from pyrhe_b200.hostmath import block_ranges
    assert block_ranges(10, 3) == [(0, 3), (3, 6), (6, 10)]
    assert block_ranges(437, 10)[-1] == (387, 437)
    assert block_ranges(8, 1) == [(0, 8)]
