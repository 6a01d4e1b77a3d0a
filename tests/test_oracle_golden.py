"""Pins oracle/rhe_oracle.py against golden vectors produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from golden_cases import CASES, SMALL_CASES
from helpers import load_golden, oracle_problem

SMALL = list(SMALL_CASES)


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)  # the goldens were made with OMP_NUM_THREADS=1
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("name", SMALL)
def test_oracle_matches_reference(name):
    from oracle import rhe_oracle
    g = load_golden(name)
    n_traits = g["T"].shape[0]
    for t in range(n_traits):
        out = rhe_oracle.run(oracle_problem(name, trait=t))
        np.testing.assert_array_equal(out["M"], g["M"])
        # same fp32 ops in the same order: agreement to fp64 round-off, far below 1e-5
        np.testing.assert_allclose(out["T"], g["T"][t], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(out["q"], g["q"][t], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(out["sigma_total"], g["res_sigma_ests_total"][t], rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(out["sigma_se"], g["res_sig_errs"][t], rtol=1e-7, atol=1e-10)
    if "XXz" in g.files:  # state arrays belong to the last trait
        for key in ("XXz", "yXXy", "UXXz", "XXUz"):
            if key in g.files:
                np.testing.assert_allclose(out[key], g[key], rtol=1e-9, atol=1e-7)


def test_oracle_decode_and_impute_bit_exact():
    from oracle import rhe_oracle
    for name in ("rhe_cov_binary", "rhe_nocov_mean", "dom_cov"):
        g = load_golden(name)
        o = rhe_oracle.Oracle(oracle_problem(name))
        blocks = [o.block_bins(j, return_block=True) for j in range(o.J)]
        got = np.concatenate(blocks, axis=1).astype(np.uint8)
        np.testing.assert_array_equal(got, g["geno_imputed"])
