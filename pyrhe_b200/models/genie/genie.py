"""GENIE: G, G+GxE and G+GxE+NxE variance components
(/root/reference/pyrhe/src/models/genie/genie.py).

GxE rows use X_gxe = diag(env) X (genie.py:65-67).  Nothing N x m is scaled: pass A sees a
second right-hand-side set env * [Z | W | y] because (diag(env) X)^T r = X^T (env * r), and pass
B scales its output rows by env.  The NxE row (X = diag(env), base.py:472-481, an N x N matrix in
the reference) is closed form: XXz = env^2 * z.

Supported: one environment column; `G+GxE+NxE` reproduces the reference
(including the Q8 quirk that the covariate terms of the NxE row use only the last random
vector); `G` and `G+GxE` crash / mislabel rows in the reference and are implemented as intended.
"""
import numpy as np

from ... import stats
from ...base import Base
from ...util.file_processing import read_env_file


class GENIE(Base):
    def __init__(self, env_file: str, genie_model: str, **kwargs):
        super().__init__(**kwargs)
        if genie_model not in ("G", "G+GxE", "G+GxE+NxE"):
            raise ValueError("Unsupported GENIE genie_model type")
        self.num_env, env = read_env_file(env_file)
        if self.num_env != 1:
            raise ValueError("GENIE supports exactly one environment column named 'env'")
        env = np.asarray(env, dtype=np.float64)
        if len(env) == self.num_indv_original and len(self.missing_indv):
            env = np.delete(env, self.missing_indv, axis=0)     # the reference forgets this filter
        self.env = env[:, np.newaxis]
        self.num_gen_env_bin = self.num_bin * self.num_env
        self.genie_model = genie_model
        self.log._log(f"Number of environments: {self.num_env}")
        self.log._log(f"GENIE model: {self.genie_model}")

    def _plan_model(self):
        return "genie"

    @property
    def _has_gxe(self):
        return self.genie_model in ("G+GxE", "G+GxE+NxE")

    @property
    def _has_nxe(self):
        return self.genie_model == "G+GxE+NxE"

    def get_num_estimates(self):
        return self.num_bin + (self.num_gen_env_bin if self._has_gxe else 0) + (self.num_env if self._has_nxe else 0)

    def get_M_last_row(self):
        parts = [self.len_bin]
        if self._has_gxe:
            parts.append(self.len_bin * self.num_env)
        if self._has_nxe:
            parts.append([1] * self.num_env)
        return np.concatenate(parts)

    def b_trace_calculation(self, k, j, b_idx):
        """genie.py:84-94: N for the G rows, the Hutchinson estimate <XXz_k, Z> / (B M_k) otherwise
        (read back from the assembled trace column)."""
        if k < self.num_bin:
            return self.num_indv
        if self._hook_mode:       # an extender's hook filled the state arrays: the reference's own expression
            return np.sum(self.XXz[k][b_idx] * self.all_zb.T) / (self.num_random_vec * self.M[j][k])
        T, _ = self.setup_lhs_rhs_jackknife(j, None)
        tr = T[k, self.num_estimates]
        if self.use_cov:   # undo the covariate correction base.py:612-618 applies on top
            ht, pc = self._host_terms, self._pieces
            plan = self._plan_cached
            G = (self._G_tot - pc["G_blk"][j] if j < self.num_jack else self._G_tot)
            H = G[k][plan.cols_W(), plan.cols_Z()] if k < plan.E_reg else ht.nxe_H
            tr = tr + np.sum(H * (ht.Q @ ht.WtZ)) / (self.num_random_vec * self.M[j][k])
        return tr

    def estimate(self, method: str = "lstsq"):
        """genie.py:97-144: also returns sigma scaled by the trace column, used for h2."""
        sigma, trace_cols = self._solve_all(method)
        adj = sigma * trace_cols
        return sigma[:-1, :], sigma[-1, :], adj[:-1, :], adj[-1, :]

    def compute_h2_nonoverlapping(self, sigma_est_jackknife, sigma_ests_total):
        """genie.py:146-188: per-component h2, then totals (all, G, GxE)."""
        s = np.vstack([sigma_est_jackknife, sigma_ests_total[np.newaxis, :]])
        K, E, G2 = self.num_bin, self.num_estimates, self.num_gen_env_bin
        denom = s[:, :E].sum(axis=1) + s[:, -1]
        h2 = s[:, :E] / denom[:, None]
        g_tot = h2[:, :K].sum(axis=1)
        gxe_tot = h2[:, K:K + G2].sum(axis=1) if self._has_gxe else np.zeros(len(s))
        nxe_tot = h2[:, K + G2:E].sum(axis=1) if self._has_nxe else np.zeros(len(s))
        cols = [h2, (g_tot + gxe_tot + nxe_tot)[:, None], g_tot[:, None]]
        if self._has_gxe:
            cols.append(gxe_tot[:, None])
        out = np.concatenate(cols, axis=1)
        return out[:-1, :], out[-1, :]

    def compute_enrichment(self, h2_jackknife, h2_total):
        """genie.py:190-219: (h2_k / M_k) / (sum h2_g / sum M) with the all-SNP bin sizes."""
        h2 = np.vstack([h2_jackknife, h2_total[np.newaxis, :]])
        K = self.num_bin
        Mk = np.asarray(self.M[-1][:K], dtype=np.float64)
        denom = h2[:, :K].sum(axis=1) / Mk.sum()
        out = (h2[:, :K] / Mk) / denom[:, None]
        return out[:-1, :], out[-1, :]

    def _component_label(self, i, prefix):
        K, G2 = self.num_bin, self.num_gen_env_bin
        if i < K:
            return f"{prefix}_g[{i}]"
        if self._has_gxe and i < K + G2:
            return f"{prefix}_gxe[{i - K}]"
        return f"{prefix}_nxe[{i - K - G2}]"

    def run(self, method):
        """Report of genie.py:221-301."""
        log = self.log
        sig_jack, sig_total, sig_jack_adj, sig_total_adj = self.estimate(method=method)
        sig_errs = self.estimate_error(sig_jack)
        E = self.num_estimates
        log._log("Variance components: ")
        for i in range(E):
            log._log(f"{self._component_label(i, 'Sigma^2')} : {sig_total[i]}  SE : {sig_errs[i]}")
        log._log(f"Sigma^2_e : {sig_total[-1]}  SE : {sig_errs[-1]}")

        h2_jack, h2_total = self.compute_h2_nonoverlapping(sig_jack_adj, sig_total_adj)
        h2_errs = self.estimate_error(h2_jack)
        log._log("*****")
        log._log("Heritabilities:")
        totals = ["Total h2", "Total h2_g"] + (["Total h2_gxe"] if self._has_gxe else [])
        for i, est in enumerate(h2_total):
            if i < E:
                log._log(f"{self._component_label(i, 'h2')} : {est} SE : {h2_errs[i]}")
            else:
                log._log(f"{totals[i - E]} : {est} SE: {h2_errs[i]}")

        log._log("*****")
        log._log("Enrichments:")
        enr_jack, enr_total = self.compute_enrichment(h2_jack, h2_total)
        enr_errs = self.estimate_error(enr_jack)
        self._log_enrichment_block(enr_total, enr_errs)
        return {"sigma_ests_total": sig_total, "sig_errs": sig_errs, "h2_total": h2_total, "h2_errs": h2_errs,
                "enrichment_total": enr_total, "enrichment_errs": enr_errs}
