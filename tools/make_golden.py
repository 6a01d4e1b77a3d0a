"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

    python tools/make_golden.py [case ...]

For each case in tools/golden_cases.py: write the synthetic data set with
pyrhe_b200.synth.make_dataset (seeded), run /root/reference's classes on it in a
fresh subprocess (PYTHONPATH=/root/reference + the bed_reader test shim,
OMP_NUM_THREADS=1, integer seed), and store T/q per jackknife, the result dict,
the log text, `.tr/.MN` text and (small cases) the state arrays and imputed
genotype counts.  /root/reference does not exist on the GPU box, so tests only
read the committed .npz files.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from golden_cases import CASES  # noqa: E402
from pyrhe_b200.synth import make_dataset  # noqa: E402

REFERENCE = "/root/reference"


def sha256_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def run_oracle_case(name, case, outdir):
    """Cases the reference cannot run (GENIE at N >= 1e5 builds an N x N matrix): vectors from the CPU oracle, which
    every reference-made golden pins.  Marked `source = "oracle"` in the file."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from oracle import rhe_oracle
    p = helpers.oracle_problem(name, verify=False)
    out = rhe_oracle.run(p)
    _, paths = helpers.case_dataset(name, verify=False)
    data = dict(T=out["T"][None], q=out["q"][None], M=out["M"], res_sigma_ests_total=out["sigma_total"][None],
                res_sig_errs=out["sigma_se"][None], sigma_jack=out["sigma_jack"][None], source=np.array("oracle"),
                bed_sha256=np.array(sha256_file(paths["geno_file"] + ".bed")),
                case=np.array(json.dumps(dict(data=case["data"], model=case["model"], kwargs=case["kwargs"]))))
    np.savez_compressed(os.path.join(outdir, name + ".npz"), **data)
    print(f"{name} (oracle): T{data['T'].shape} sigma={out['sigma_total']}")


def run_case(name, case, outdir):
    if case.get("source") == "oracle":
        return run_oracle_case(name, case, outdir)
    tmp = tempfile.mkdtemp(prefix=f"golden_{name}_")
    paths = make_dataset(tmp, name, **case["data"])
    spec = dict(model=case["model"], kwargs=case["kwargs"], paths=paths,
                dump_state=case.get("dump_state", False))
    spec_path = os.path.join(tmp, "spec.json")
    json.dump(spec, open(spec_path, "w"))
    raw_out = os.path.join(tmp, "ref.npz")
    env = dict(os.environ)
    env["PYTHONPATH"] = REFERENCE + os.pathsep + os.path.join(ROOT, "tests", "_shim")
    env["OMP_NUM_THREADS"] = "1"
    env["MKL_NUM_THREADS"] = "1"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "_ref_worker.py"), spec_path, raw_out],
                   check=True, env=env, cwd=tmp)
    data = dict(np.load(raw_out, allow_pickle=False))
    data["bed_sha256"] = np.array(sha256_file(paths["geno_file"] + ".bed"))
    data["case"] = np.array(json.dumps(dict(data=case["data"], model=case["model"], kwargs=case["kwargs"])))
    np.savez_compressed(os.path.join(outdir, name + ".npz"), **data)
    print(f"{name}: T{data['T'].shape} sigma={data['res_sigma_ests_total'][0]}")


def main():
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    names = sys.argv[1:] or list(CASES)
    for name in names:
        run_case(name, CASES[name], outdir)


if __name__ == "__main__":
    main()
