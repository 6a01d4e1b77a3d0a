"""Child process of oracle/ref_baseline.py: the UNMODIFIED reference (oracle/_ref/pyrhe on PYTHONPATH) on one sample.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Usage: _ref_run.py spec.json -- prints one JSON line with the seconds of
`model(trait=0)` and, inside it, of `pre_compute` (every jackknife block through read_geno / impute / bins / standardise /
XXz, UXXz, XXUz, yXXy + aggregate: the hot path; the rest of the call is the J + 1 normal equations, solves and h2)."""
import io
import json
import sys
import time


def main():
    spec = json.load(open(sys.argv[1]))
    import torch
    torch.set_num_threads(spec["threads"])
    from pyrhe.src.models.rhe import RHE
    from pyrhe.src.util.logger import Logger
    log = Logger(suppress=True, debug_mode=False)
    real_stdout = sys.stdout
    sys.stdout = io.StringIO()
    t0 = time.perf_counter()
    model = RHE(model="rhe", log=log, multiprocessing=spec["workers"] > 1, device="cpu", num_workers=spec["workers"],
                **spec["kwargs"], **spec["paths"])
    t1 = time.perf_counter()
    out = {"constructor_s": t1 - t0, "num_indv": int(model.num_indv), "num_snp": int(model.num_snp)}
    if spec.get("full_call"):
        # the whole user call: pre_compute, then run() = the J + 1 normal equations (O(J E^2 B N) on the host), solves, h2
        timing = {}
        inner = model.pre_compute

        def timed_pre_compute(*a, **k):                # a stopwatch around the reference's own method, nothing inside it is touched
            ta = time.perf_counter()
            r = inner(*a, **k)
            timing["pre_compute_s"] = time.perf_counter() - ta
            return r

        model.pre_compute = timed_pre_compute
        res = model(trait=0)
        out.update(call_s=time.perf_counter() - t1, pre_compute_s=timing.get("pre_compute_s"),
                   sigma_e=float(res["sigma_ests_total"][-1]))
    else:
        # the hot path alone, entered the way Base.__call__ enters it (base.py:876,879)
        model.pheno = model.pheno_cp[:, 0].reshape(-1, 1)
        ta = time.perf_counter()
        model.pre_compute()
        out.update(pre_compute_s=time.perf_counter() - ta, call_s=None)
    sys.stdout = real_stdout
    print(json.dumps(out))


if __name__ == "__main__":
    main()
