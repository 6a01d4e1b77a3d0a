"""Definitions of the golden parity cases (shared by tools/make_golden.py and tests/).

Every case is (dataset spec for pyrhe_b200.synth.make_dataset, model class name,
constructor kwargs).  Inputs are regenerated deterministically from the seed at
test time; the golden file stores a sha256 of the `.bed` so a drifted generator
is caught instead of silently comparing different data.
"""

CASES = {
    # RHE, covariates, binary imputation with missing genotypes, dropped individuals,
    # N0 not a multiple of 4, M not a multiple of J, two traits.
    "rhe_cov_binary": dict(
        data=dict(N=203, M=437, K=3, seed=11, n_cov=3, n_traits=2, missing_rate=0.01,
                  missing_pheno=(5, 77)),
        model="RHE", kwargs=dict(num_jack=10, num_random_vec=6, geno_impute_method="binary", seed=7),
        dump_state=True),
    # RHE, single bin, no covariates, mean imputation.
    "rhe_nocov_mean": dict(
        data=dict(N=200, M=400, K=1, seed=12, missing_rate=0.02),
        model="RHE", kwargs=dict(num_jack=8, num_random_vec=5, geno_impute_method="mean", seed=3),
        dump_state=True),
    # RHE with an overlapping annotation (a SNP may sit in two bins).
    "rhe_overlap": dict(
        data=dict(N=301, M=512, K=4, seed=13, overlap=0.3),
        model="RHE", kwargs=dict(num_jack=6, num_random_vec=4, geno_impute_method="binary", seed=5),
        dump_state=False),
    # One jackknife block (the `num_jack == 1` branch of Base.estimate).
    "rhe_one_block": dict(
        data=dict(N=128, M=256, K=2, seed=14),
        model="RHE", kwargs=dict(num_jack=1, num_random_vec=3, geno_impute_method="mean", seed=1),
        dump_state=False),
    # Additive + dominance, with and without covariates.
    "dom_cov": dict(
        data=dict(N=222, M=410, K=2, seed=21, n_cov=2, missing_rate=0.01),
        model="RHE_DOM", kwargs=dict(num_jack=7, num_random_vec=5, geno_impute_method="binary", seed=9),
        dump_state=True),
    "dom_nocov": dict(
        data=dict(N=180, M=360, K=3, seed=22),
        model="RHE_DOM", kwargs=dict(num_jack=6, num_random_vec=4, geno_impute_method="mean", seed=2),
        dump_state=False),
    # GENIE.
    "genie_full_cov": dict(
        data=dict(N=210, M=420, K=2, seed=31, n_cov=2, with_env=True, missing_rate=0.005),
        model="GENIE", kwargs=dict(num_jack=7, num_random_vec=5, geno_impute_method="binary", seed=4,
                                   genie_model="G+GxE+NxE"),
        dump_state=True),
    "genie_full_nocov": dict(
        data=dict(N=190, M=380, K=1, seed=32, with_env=True),
        model="GENIE", kwargs=dict(num_jack=5, num_random_vec=6, geno_impute_method="mean", seed=6,
                                   genie_model="G+GxE+NxE"),
        dump_state=False),
    # (genie_model "G" and "G+GxE" crash / mislabel rows in the reference itself -- base.py:467-474,
    #  SURVEY.md Q7 -- so they have no golden; the product implements them as intended.)
    # Shape of the reference's example config (N=5000, M=10000, 8 bins, 5 cov, B=10, J=100).
    "rhe_example_shape": dict(
        data=dict(N=5000, M=10000, K=8, seed=41, n_cov=5),
        model="RHE", kwargs=dict(num_jack=100, num_random_vec=10, geno_impute_method="binary", seed=0),
        dump_state=False),
}
