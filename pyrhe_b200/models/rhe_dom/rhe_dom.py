"""RHE_DOM: additive + dominance components, 2 K estimates
(/root/reference/pyrhe/src/models/rhe_dom/rhe_dom.py).

The dominance operand h = mu g - 2 [g == 2] (rhe_dom.py:23-41 with maf = mu / 2) is never
materialised: pass A multiplies the packed bytes once as counts and once as the [g == 2]
indicator, and the per-SNP scaling 1 / (mu (1 - mu/2)) is applied to the outputs."""
import numpy as np

from ...base import Base


class RHE_DOM(Base):
    def _plan_model(self):
        return "rhe_dom"

    def get_num_estimates(self):
        return self.num_bin * 2

    def get_M_last_row(self):
        return np.concatenate([self.len_bin, self.len_bin])

    def b_trace_calculation(self, k, j, b_idx):
        return self.num_indv

    def run(self, method):
        """Report of rhe_dom.py:70-117."""
        log = self.log
        sig_jack, sig_total = self.estimate(method=method)
        sig_errs = self.estimate_error(sig_jack)
        self._log_variance_components(sig_total, sig_errs)
        h2_jack, h2_total = self.compute_h2_nonoverlapping(sig_jack, sig_total)
        h2_errs = self.estimate_error(h2_jack)
        log._log("*****")
        log._log("Heritabilities:")
        self._log_h2_block(h2_total, h2_errs)
        log._log("*****")
        log._log("Enrichments: ")
        enr_jack, enr_total = self.compute_enrichment(h2_jack, h2_total)
        enr_errs = self.estimate_error(enr_jack)
        self._log_enrichment_block(enr_total, enr_errs)
        return {"sigma_ests_total": sig_total, "sig_errs": sig_errs, "h2_total": h2_total, "h2_errs": h2_errs,
                "enrichment_total": enr_total, "enrichment_errs": enr_errs}
