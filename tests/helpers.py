"""Shared test plumbing: regenerate a golden case's inputs and hand them to the oracle."""
import hashlib
import json
import os
import tempfile

import numpy as np

from golden_cases import CASES
from pyrhe_b200.synth import make_dataset
from pyrhe_b200.util import file_processing as fp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
_CACHE = {}


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def case_dataset(name, verify=True):
    """Regenerate the inputs of a golden case (cached per session); verifies the .bed hash."""
    if name in _CACHE:
        return _CACHE[name]
    case = CASES[name]
    tmp = tempfile.mkdtemp(prefix=f"case_{name}_")
    paths = make_dataset(tmp, name, **case["data"])
    if verify:
        g = load_golden(name)
        with open(paths["geno_file"] + ".bed", "rb") as f:
            digest = hashlib.sha256(f.read()).hexdigest()
        assert digest == str(g["bed_sha256"]), "synthetic generator drifted from the golden inputs"
    _CACHE[name] = (case, paths)
    return case, paths


def oracle_problem(name, trait=0, verify=True):
    """Parse the case's files with the product's text front end and build an OracleProblem."""
    from oracle.rhe_oracle import OracleProblem
    case, paths = case_dataset(name, verify)
    kw = case["kwargs"]
    N0, _ = fp.read_fam(paths["geno_file"] + ".fam")
    M = fp.read_bim(paths["geno_file"] + ".bim")
    _, annot, _ = fp.read_annot(paths["annot_file"], kw["num_jack"])
    y, missing, _ = fp.read_pheno(paths["pheno_file"])
    W = None
    if "cov_file" in paths:
        W, missing = fp.read_cov(paths["cov_file"], missing_indvs=missing)
    y = np.delete(y, missing, axis=0)
    y = y - np.mean(y, axis=0)
    N = N0 - len(missing)
    np.random.seed(kw["seed"])
    Z = np.random.randn(N, kw["num_random_vec"])
    packed = np.fromfile(paths["geno_file"] + ".bed", dtype=np.uint8, offset=3).reshape(M, (N0 + 3) // 4)
    env = None
    if "env_file" in paths:
        env = fp.read_env_file(paths["env_file"])[1]
    model = {"RHE": "rhe", "RHE_DOM": "rhe_dom", "GENIE": "genie"}[case["model"]]
    return OracleProblem(packed=packed, n_indv_original=N0, annot=annot, Z=Z, y=y[:, trait].reshape(-1, 1),
                         num_jack=kw["num_jack"], W=W, missing_indv=tuple(missing),
                         impute=kw["geno_impute_method"], seed=kw["seed"], model=model,
                         genie_model=kw.get("genie_model", "G+GxE+NxE"), env=env)
