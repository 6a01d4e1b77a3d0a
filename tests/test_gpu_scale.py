"""GPU: parity at BASELINE scale in N (VERDICT r1 "pin parity at BASELINE scale").

The unmodified reference ran these cases in this repo's build container (tools/make_golden.py; N = 500k / 200k / 100k /
20k individuals, M = 800 SNPs in 4 jackknife blocks, 8 bins, 5 covariates, B = 10, missing genotypes with binary
imputation); the GENIE case at N = 300k comes from the CPU oracle because the reference's NxE row needs an N x N
matrix.  The CUDA path (int8 tcgen05 kernels, through the C ABI) is held to them for every quantity north_star names:
per-jackknife T and q, variance components, h2, jackknife SEs, and -- at model level -- enrichments.

Tolerances (the ones SURVEY.md §9.2 derived from the reference's own fp32-vs-fp64 gap):
    T, q          rtol 1e-5, atol 1e-6 * max|.|
    sigma^2, h2   rtol 1e-5, atol 1e-7 * Var(y)   (h2: 1e-7)
    SE            rtol 1e-4, atol 1e-7 * Var(y)
The measured worst-case errors of every case go to `parity_report.json` (gpurun_out/ on the GPU box; the committed
copy is tests/parity_report.json).
"""
import json
import os

import numpy as np
import pytest

from golden_cases import CASES, SCALE_CASES
from helpers import ROOT, case_dataset, load_golden, oracle_problem
from device_model import assemble_all
from test_gpu_parity import make_engine, plan_for, solve_all

pytestmark = pytest.mark.gpu
REPORT = {}


def _rel(got, ref, floor):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), floor)))


def _record(name, entry):
    REPORT[name] = entry
    for d in (os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "tests")):
        try:
            os.makedirs(d, exist_ok=True)
            path = os.path.join(d, "parity_report.json")
            old = json.load(open(path)) if os.path.exists(path) else {}
            old.update(REPORT)
            json.dump(old, open(path, "w"), indent=1, sort_keys=True)
        except OSError:
            pass


@pytest.mark.parametrize("name", SCALE_CASES)
def test_scale_parity_engine(name):
    g = load_golden(name)
    p = oracle_problem(name)
    plan = plan_for(p)
    eng, ht, _ = make_engine(p, plan, kernel_path=1)
    pieces = eng.run()
    launches = eng.launches
    eng.close()
    J = p.num_jack
    np.testing.assert_array_equal(pieces["M"], g["M"])
    T, q = assemble_all(plan, ht, pieces, J)
    Tr, qr = g["T"][0], g["q"][0]
    sig = solve_all(T, q)
    sig_ref = solve_all(Tr, qr)
    vy = float(np.var(p.y))
    se = np.sqrt((J - 1) * ((sig[:-1] - sig[:-1].mean(0)) ** 2).sum(0) / J)
    se_ref = np.asarray(g["res_sig_errs"][0])
    tot_ref = np.asarray(g["res_sigma_ests_total"][0])
    np.testing.assert_allclose(sig_ref[-1], tot_ref, rtol=1e-9, atol=1e-12)       # the golden is self-consistent
    E = plan.E
    h2 = sig[:, :E] / sig[:, : E + 1].sum(1, keepdims=True)
    h2_ref = sig_ref[:, :E] / sig_ref[:, : E + 1].sum(1, keepdims=True)
    entry = {
        "source": str(g["source"]) if "source" in g.files else "unmodified reference (tools/make_golden.py)",
        "N": int(p.Z.shape[0]), "M": int(p.annot.shape[0]), "J": J, "E": E, "kernel_path": "tcgen05",
        "kernel_launches": int(launches), "var_y": vy,
        "T_max_rel(floor 1e-6*max)": _rel(T, Tr, 1e-6 * np.abs(Tr).max()),
        "q_max_rel(floor 1e-6*max)": _rel(q, qr, 1e-6 * np.abs(qr).max()),
        "sigma2_total_max_abs/var_y": float(np.max(np.abs(sig[-1] - tot_ref)) / vy),
        "sigma2_total_max_rel(floor 1e-7*var_y)": _rel(sig[-1], tot_ref, 1e-7 * vy),
        "sigma2_jackknife_max_rel(floor 1e-7*var_y)": _rel(sig[:-1], sig_ref[:-1], 1e-7 * vy),
        "h2_max_abs": float(np.max(np.abs(h2 - h2_ref))),
        "se_max_rel(floor 1e-7*var_y)": _rel(se, se_ref, 1e-7 * vy),
    }
    _record(name, entry)
    np.testing.assert_allclose(T, Tr, rtol=1e-5, atol=1e-6 * np.abs(Tr).max())
    np.testing.assert_allclose(q, qr, rtol=1e-5, atol=1e-6 * np.abs(qr).max())
    np.testing.assert_allclose(sig[-1], tot_ref, rtol=1e-5, atol=1e-7 * vy)
    np.testing.assert_allclose(sig[:-1], sig_ref[:-1], rtol=1e-5, atol=1e-7 * vy)
    np.testing.assert_allclose(h2, h2_ref, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(se, se_ref, rtol=1e-4, atol=1e-7 * vy)


@pytest.mark.parametrize("name,streaming", [("scale_rhe_500k", True), ("scale_dom_200k", True), ("scale_genie_20k", False)])
def test_scale_parity_model_api(name, streaming, tmp_path):
    """The drop-in classes end to end at scale (files -> ingest ring -> kernels -> result dict): every entry of the
    reference's result dict, including enrichments and the overlapping-annotation h2."""
    import pyrhe.models as models
    from pyrhe.src.util import Logger
    g = load_golden(name)
    case, paths = case_dataset(name)
    cls = getattr(models, ("Streaming" if streaming else "") + case["model"])
    kw = dict(case["kwargs"])
    kw.update(paths)
    model = cls(model=case["model"].lower(), log=Logger(suppress=True, debug_mode=False), multiprocessing=False,
                device="cuda", num_workers=1, **kw)
    res = model(trait=0)
    vy = float(np.var(model.pheno_cp[:, 0]))
    entry = {}
    for key, val in res.items():
        ref = np.asarray(g["res_" + key][0], dtype=np.float64)
        val = np.asarray(val, dtype=np.float64)
        is_se = "err" in key
        floor = 1e-7 * (vy if "sig" in key else 1.0)
        entry[key + "_max_rel"] = _rel(val, ref, floor)
        np.testing.assert_allclose(val, ref, rtol=1e-4 if is_se else 1e-5, atol=floor, err_msg=f"{name} {key}")
    entry["class"] = cls.__name__
    entry["ingest"] = getattr(model, "ingest_report", None)
    _record(name + ":model_api", entry)
