"""Text-format front end: `.fam/.bim`, annotation, phenotype, covariate and environment files.

Format-compatible with /root/reference/pyrhe/src/util/file_processing.py (same
function names, arguments and return values) -- SURVEY.md §2 row 9 / §9.4.  This is
O(N+M) host work and is not accelerated.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from .types import *  # noqa: F401,F403


def read_bim(filename):
    """Number of SNPs = number of lines (file_processing.py:6-23 counts '#' lines too)."""
    try:
        with open(filename, "r") as fh:
            return sum(1 for _ in fh)
    except FileNotFoundError:
        raise FileNotFoundError(f"Error: The bim file {filename} could not be found.")


def read_fam(filename):
    """(number of individuals, DataFrame of the .fam columns) -- file_processing.py:25-35."""
    try:
        df = pd.read_csv(filename, sep=r"\s+", header=None)
    except FileNotFoundError:
        raise FileNotFoundError(f"Error: The fam file '{filename}' could not be found.")
    return df.shape[0], df


def read_annot(filename, Njack):
    """(K, annot [M, K] int, len_bin [K]) -- file_processing.py:37-69."""
    rows = []
    try:
        with open(filename, "r") as fh:
            for line in fh:
                if line.startswith("#"):
                    continue
                rows.append([int(tok) for tok in line.split()])
    except FileNotFoundError:
        raise FileNotFoundError("Error: The annotation file could not be found.")
    annot = np.array(rows)
    n_bin = annot.shape[1] if annot.ndim == 2 else 0
    len_bin = (annot == 1).sum(axis=0).astype(int) if n_bin else np.zeros(0, dtype=int)
    return n_bin, annot, len_bin


def read_pheno(filename):
    """(y [N0, T], rows with any NA/-9, all-values-in-{0,1,2}) -- file_processing.py:72-107."""
    try:
        with open(filename, "r") as fh:
            lines = fh.readlines()
    except FileNotFoundError:
        raise FileNotFoundError("Error: The pheno file could not be found.")
    n_pheno = len(lines[0].split()) - 2
    y, missing, all_binary = [], [], True
    for i, line in enumerate(lines[1:]):
        toks = line.split()[2:]
        vals = [-9.0 if t == "NA" else float(t) for t in toks]
        if "NA" in toks or -9 in vals:
            missing.append(i)
            y.append([-9] * n_pheno)
            continue
        if any(v not in (0, 1, 2) for v in vals):
            all_binary = False
        y.append(vals)
    return np.array(y), missing, all_binary


def generate_annot(filename, num_snp, num_bin):
    """Random one-bin-per-SNP annotation drawn from numpy's GLOBAL RNG, one scalar draw per
    SNP (file_processing.py:109-118) -- it runs between `np.random.seed` and the draw of Z
    (base.py:73,112,176), so the draw order is part of the parity contract."""
    with open(filename, "w") as fh:
        for _ in range(num_snp):
            row = ["0"] * num_bin
            row[np.random.randint(0, num_bin)] = "1"
            fh.write(" ".join(row) + "\n")


def read_cov(filename, std: bool = False, missing_indvs: list = None, cov_impute_method: str = "ignore",
             one_hot_conversion: bool = False, categorical_threshold: int = 100, logger=None):
    """(covariate matrix, all missing rows) -- file_processing.py:121-199.

    `one_hot_conversion` only writes `<column>_one_hot.cov` side files; the returned matrix
    keeps the original columns.

    `cov_impute_method="mean"`: NA / -9 cells are replaced by their column mean and the individual is KEPT.  The
    reference reports those individuals as missing all the same (file_processing.py:157) while their rows stay in the
    covariate matrix, so phenotype and covariates end up with different lengths (and under pandas >= 3 its chained
    `fillna(inplace=True)` no longer fills at all): it cannot run such a file.  This is the intended behaviour."""
    try:
        df = pd.read_csv(filename, sep=r"\s+")
    except FileNotFoundError:
        raise FileNotFoundError(f"Error: The covariate file '{filename}' could not be found.")
    missing_indvs = list(missing_indvs) if missing_indvs else []
    if missing_indvs:
        df = df.drop(index=missing_indvs, errors="ignore")
    df = df.drop(columns=[c for c in ("FID", "IID") if c in df.columns])
    is_missing = df.replace("NA", np.nan).isin([np.nan, -9]).any(axis=1)
    newly_missing = df.index[is_missing].tolist()
    if cov_impute_method == "ignore":
        df = df[~is_missing]
    else:
        df = df.replace({"NA": np.nan, "-9": np.nan, -9: np.nan}).astype(float)
        df = df.fillna(df.mean())
        newly_missing = []
    for column in df.columns:
        n_unique = df[column].nunique()
        if n_unique <= categorical_threshold:
            if one_hot_conversion:
                if logger:
                    logger._debug(f"Column '{column}' detected as categorical with {n_unique} unique values.")
                one_hot = pd.get_dummies(df[column], prefix=column, drop_first=False).astype(int)
                side_file = f"{column}_one_hot.cov"
                one_hot.to_csv(side_file, index=False, sep=" ", header=False)
                if logger:
                    logger._debug(f"One-hot encoded values for '{column}' stored in '{side_file}'")
        elif logger:
            logger._debug(f"Column '{column}' contains quantitative values (number of unique values is "
                          f"{n_unique} while categorical threshold is {categorical_threshold})")
    if std:
        df = (df - df.mean()) / df.std(ddof=1)
    return df.values, missing_indvs + newly_missing


def read_env_file(file_path):
    """(number of environments, the column named `env`) -- file_processing.py:212-227."""
    try:
        df = pd.read_csv(file_path, sep=r"\s+")
    except FileNotFoundError:
        raise FileNotFoundError(f"Error: The file '{file_path}' could not be found.")
    return len(df.columns) - 2, df["env"].to_numpy()
