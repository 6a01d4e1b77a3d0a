"""Runs ONE golden case through the UNMODIFIED reference (fresh process, pinned threads).

Invoked by tools/make_golden.py with PYTHONPATH=/root/reference:tests/_shim and
OMP_NUM_THREADS=1 (a process that has run torch ops hangs on
fork, and thread counts change fp32 sums).  Usage: _ref_worker.py spec.json out.npz
"""
import io
import json
import os
import sys
import tempfile

import numpy as np


def main():
    spec = json.load(open(sys.argv[1]))
    out_path = sys.argv[2]
    from pyrhe.src.models.rhe import RHE, StreamingRHE  # noqa: F401
    from pyrhe.src.models.rhe_dom import RHE_DOM, StreamingRHE_DOM  # noqa: F401
    from pyrhe.src.models.genie import GENIE, StreamingGENIE  # noqa: F401
    from pyrhe.src.util.logger import Logger

    cls = {"RHE": RHE, "RHE_DOM": RHE_DOM, "GENIE": GENIE}[spec["model"]]
    log = Logger(suppress=True, debug_mode=False)
    trace_dir = tempfile.mkdtemp()
    kwargs = dict(spec["kwargs"])
    kwargs.update(spec["paths"])
    model_name = {"RHE": "rhe", "RHE_DOM": "rhe_dom", "GENIE": "genie"}[spec["model"]]
    real_stdout = sys.stdout
    sys.stdout = io.StringIO()  # GENIE prints per-bin lines (genie.py:50)
    model = cls(model=model_name, log=log, multiprocessing=False, device="cpu", num_workers=1,
                get_trace=True, trace_dir=trace_dir, **kwargs)

    if getattr(model, "binary_pheno", False):
        # run() calls a method name that does not exist (rhe.py:84-87), and hands the total row the
        # whole h2 / SE lists instead of its own entry.  Alias the method that exists and pick the total's entry: the
        # arithmetic (base.py:857-868) is the reference's own.
        def _liability(h2, se):
            if np.ndim(h2):
                h2, se = h2[-1], se[-1]
            return model._compute_liability_h2(h2, se)
        model.calculate_liability_h2 = _liability

    Ts, qs = [], []
    orig = model.setup_lhs_rhs_jackknife

    def capture(j, trace_sums, is_streaming=False):
        T, q = orig(j, trace_sums, is_streaming)
        Ts.append(np.array(T, dtype=np.float64))
        qs.append(np.array(q, dtype=np.float64).ravel())
        return T, q

    model.setup_lhs_rhs_jackknife = capture

    res_all = {}
    for t in range(model.num_traits):
        res = model(trait=t)
        for k, v in res.items():
            res_all.setdefault(k, []).append(np.asarray(v, dtype=np.float64))
    sys.stdout = real_stdout

    J = model.num_jack
    nT = model.num_traits
    out = {
        "T": np.array(Ts).reshape(nT, J + 1, *Ts[0].shape),
        "q": np.array(qs).reshape(nT, J + 1, -1),
        "M": np.asarray(model.M, dtype=np.int64),
        # tests regenerate Z from the seed (base.py:73,176); the stored copy is a cross-check for the small cases only
        "Z": np.asarray(model.all_zb if model.all_zb.size <= 100_000 else model.all_zb[:1000], dtype=np.float64),
        "missing_indv": np.asarray(model.missing_indv, dtype=np.int64),
        "num_indv": np.int64(model.num_indv),
        "log": np.array("".join(log.msgs)),
    }
    for k, v in res_all.items():
        out["res_" + k] = np.array(v)
    stem = os.path.join(trace_dir, "run_" + os.path.basename(spec["paths"]["pheno_file"]))
    out["tr_text"] = np.array(open(stem + ".tr").read())
    out["mn_text"] = np.array(open(stem + ".MN").read())

    if spec.get("dump_state"):
        # state arrays after aggregate() of the LAST trait (LOO sums; slot J = totals)
        out["XXz"] = np.asarray(model.XXz, dtype=np.float64)
        out["yXXy"] = np.asarray(model.yXXy, dtype=np.float64)
        if model.use_cov:
            out["UXXz"] = np.asarray(model.UXXz, dtype=np.float64)
            out["XXUz"] = np.asarray(model.XXUz, dtype=np.float64)
        # decoded + imputed genotype counts, block by block, exactly as _pre_compute_worker does
        cols = []
        raw_cols = []
        for j in range(J):
            np.random.seed(model.seed)
            sub, _ = model._get_jacknife_subsample(j)
            raw = sub.copy()
            raw[np.isnan(raw)] = 3
            raw_cols.append(raw.astype(np.uint8))
            cols.append(model.impute_geno(sub).astype(np.uint8))
        out["geno_imputed"] = np.concatenate(cols, axis=1)  # [N, M]
        out["geno_raw"] = np.concatenate(raw_cols, axis=1)  # [N, M], 3 = missing
    np.savez_compressed(out_path, **out)


if __name__ == "__main__":
    main()
