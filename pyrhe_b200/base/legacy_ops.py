"""Per-op helpers of the reference's `Base` kept for custom subclasses ("extenders").

The built-in models never call these: their whole block loop is fused in libpyrhe_b200.  A subclass written
against /root/reference/pyrhe/src/base/base.py that overrides `pre_compute_jackknife_bin` and calls
`read_geno` / `impute_geno` / `partition_bins` / `standardize_geno` / `_compute_XXz` ... still finds them here,
with the same signatures and return types (numpy in, numpy out).  `read_geno` decodes on the GPU with the
library's decode kernel; the dense products go through `mat_mul` (torch on CUDA, fp32), as in the reference.
There is no CPU implementation behind any of them.
"""
from __future__ import annotations

import numpy as np

from ..util.mat_mul import mat_mul, to_tensor


class LegacyBlockOps:
    # ---- a1: .bed decode (base.py:338-359)
    def _decode_engine(self):
        eng = getattr(self, "_dec_engine", None)
        if eng is None:
            from ..assemble import PathPlan
            from ..engine import RheEngine
            from .. import _lib
            keep = np.ones(self.num_indv_original, dtype=bool)
            annot = np.ones((self.num_snp, 1), dtype=np.int64)
            eng = RheEngine(PathPlan(model="rhe", K=1, B=1, C=0), n_indv=self.num_indv_original, keep=keep, annot=annot,
                            num_jack=1, impute="mean", seed=0, device=self.device, kernel_path=_lib.PATH_SIMT)
            zero = np.zeros((self.num_indv_original, 1))
            eng.set_rhs(zero, None, zero)
            self._dec_engine = eng
        return eng

    def read_geno(self, start, end):
        """float32 [N, end - start], value = count of the second .bim allele, NaN = missing; rows of
        individuals with missing phenotype / covariates removed."""
        try:
            counts = self._decode_engine().decode_rows(np.ascontiguousarray(self.geno_bed[start:end]))  # [m, N0] int8
            geno = counts.T.astype(np.float32)
            geno[geno == 3] = np.nan
            if len(self.missing_indv):
                geno = np.delete(geno, self.missing_indv, axis=0)
            return geno
        except Exception as e:
            raise Exception(f"Error occurred: {e}")

    def _get_jacknife_subsample(self, jack_index: int):
        """(genotypes, annotation rows) of jackknife block `jack_index` -- base.py:362-379."""
        from ..hostmath import block_ranges
        start, end = block_ranges(self.num_snp, self.num_jack)[jack_index]
        return self.read_geno(start, end), self.annot_matrix[start:end]

    def _get_annot_subsample(self, jack_index: int):
        """Annotation with the block's rows removed (base.py:382-393, including its j == num_jack quirk)."""
        step = self.num_snp // self.num_jack
        start = jack_index * step
        end = start + (step if jack_index < self.num_jack - 1 else step + self.num_snp % self.num_jack)
        mask = np.ones(self.num_snp, dtype=bool)
        mask[start:end] = False
        return self.annot_matrix[mask]

    # ---- a2: imputation (base.py:265-289)
    def _simulate_geno_from_random(self, p_j):
        rval = np.random.random()
        d0, d1 = (1 - p_j) * (1 - p_j), 2 * p_j * (1 - p_j)
        return 0 if rval < d0 else (1 if rval < d0 + d1 else 2)

    def impute_geno(self, X):
        """In place; "binary" draws ONE uniform per SNP from numpy's global RNG (even without missing entries),
        "mean" fills 0 (base.py:277-289)."""
        import torch
        Xt = to_tensor(np.ascontiguousarray(X), self.device)
        miss = torch.isnan(Xt)
        if self.geno_impute_methods == "binary":
            observed = torch.nanmean(Xt, dim=0).cpu().numpy()
            fills = np.array([self._simulate_geno_from_random(np.float32(m) * 0.5) for m in observed], dtype=np.float32)
            Xt = torch.where(miss, torch.from_numpy(fills).to(Xt.device)[None, :], Xt)
        else:
            Xt = torch.where(miss, torch.zeros((), device=Xt.device), Xt)
        X[...] = Xt.cpu().numpy()
        return X

    # ---- a3: bin gather (base.py:315-336)
    def _bin_to_snp(self, annot):
        return [np.nonzero(annot[:, k])[0].tolist() for k in range(self.num_bin)]

    def partition_bins(self, geno: np.ndarray, annot: np.ndarray):
        return [geno[:, idx] for idx in self._bin_to_snp(annot)]

    # ---- a4: standardisation (base.py:291-296)
    def standardize_geno(self, geno):
        import torch
        g = to_tensor(np.ascontiguousarray(geno), self.device)
        mu = g.mean(dim=0)
        return ((g - mu) * torch.rsqrt(mu * (1 - 0.5 * mu))).cpu().numpy()

    # ---- a5-a8: the four statistics (base.py:403-417)
    def _compute_XXz(self, b, X_kj):
        z = self.all_zb[:, b].reshape(-1, 1)
        return mat_mul(X_kj, mat_mul(X_kj.T, z, device=self.device), device=self.device).flatten()

    def _compute_UXXz(self, XXz_kjb):
        W = self.cov_matrix
        return mat_mul(W, mat_mul(self.Q, mat_mul(W.T, XXz_kjb, device=self.device), device=self.device),
                       device=self.device).flatten()

    def _compute_XXUz(self, b, X_kj):
        uz = self.all_Uzb[:, b].reshape(-1, 1)
        return mat_mul(X_kj, mat_mul(X_kj.T, uz, device=self.device), device=self.device).flatten()

    def _compute_yXXy(self, X_kj, y):
        pheno = y if not self.use_cov else self.regress_pheno(self.cov_matrix, y)
        v = mat_mul(X_kj.T, pheno, device=self.device)
        return mat_mul(v.T, v, device=self.device)
