"""RHE: one variance component per annotation bin (/root/reference/pyrhe/src/models/rhe/rhe.py)."""
from ...base import Base


class RHE(Base):
    def get_num_estimates(self):
        return self.num_bin

    def get_M_last_row(self):
        return self.len_bin

    def b_trace_calculation(self, k, j, b_idx):
        # tr(X X^T)/M = N for standardised genotypes (rhe.py:24-26)
        return self.num_indv

    def run(self, method):
        """Report text and result dict of rhe.py:28-101 (log lines are parsed by downstream tests)."""
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

        log._log("*****\n*****\nHeritabilities and enrichments computed based on overlapping setting")
        h2o_jack, h2o_total = self.compute_h2_overlapping(sig_jack, sig_total)
        h2o_errs = self.estimate_error(h2o_jack)
        log._log("Heritabilities:")
        self._log_h2_block(h2o_total, h2o_errs)
        log._log("Enrichments (overlapping def):")
        enro_jack, enro_total = self.compute_enrichment(h2o_jack, h2o_total)
        enro_errs = self.estimate_error(enro_jack)
        self._log_enrichment_block(enro_total, enro_errs)

        if self.binary_pheno and self.samp_prev is not None and self.pop_prev is not None:
            log._log("*****")
            log._log("Liability Scale h2 for binary phenotype:")
            last = len(h2_total) - 1
            for i, est in enumerate(h2_total):
                h, se, pval = self._compute_liability_h2(est, h2_errs[i])
                if i == last:
                    log._log(f"Total Liability-scale h2 : {h}, SE: {se}, p-value: {pval}")
                else:
                    log._log(f"Liability-scale h2_g[{i}] : {h}, SE: {se}, p-value: {pval}")

        return {
            "sigma_ests_total": sig_total, "sig_errs": sig_errs,
            "h2_total": h2_total, "h2_errs": h2_errs,
            "enrichment_total": enr_total, "enrichment_errs": enr_errs,
            "h2_total_overlap": h2o_total, "h2_errs_overlap": h2o_errs,
            "enrichment_total_overlap": enro_total, "enrichment_errs_overlap": enro_errs,
        }
