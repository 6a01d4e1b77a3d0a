"""`Base`: the reference's model framework class, re-hosted on the fused B200 block engine.

Public surface kept from /root/reference/pyrhe/src/base/base.py:
constructor kwargs (:24-48), attributes (`num_traits`, `num_indv`, `num_snp`, `num_bin`,
`len_bin`, `annot_matrix`, `all_zb`, `all_Uzb`, `M`, ...), the abstract block interface
(`get_num_estimates`, `get_M_last_row`, `pre_compute_jackknife_bin`, `b_trace_calculation`,
`run`), `pre_compute` / `aggregate` / `setup_lhs_rhs_jackknife` / `estimate` /
`estimate_error` / `compute_h2_*` / `compute_enrichment` / `get_trace_summary`, and
`model(trait, method="QR") -> dict`.

What changed underneath: `pre_compute` no longer forks workers that call
`pre_compute_jackknife_bin` ~58k times through `mat_mul` (base.py:503-554); it makes one
pass over the packed `.bed` with libpyrhe_b200 (engine.py) and keeps only small Gram
pieces on the host.  `num_workers` / `multiprocessing` are accepted for compatibility;
data parallelism is one process per GPU under torch.distributed.
"""
from __future__ import annotations

import os
import time
from abc import ABC, abstractmethod
from typing import List, Optional, Tuple

import numpy as np

from .. import stats
from ..assemble import (PathPlan, loo_grams, normal_equations, normal_equations_batch, normal_equations_finish,
                        normal_equations_prepare, trace_sums_row)
from ..hostmath import block_ranges, host_terms
from ..util.file_processing import (generate_annot, read_annot, read_bim, read_cov, read_fam, read_pheno)
from ..util.logger import Logger
from ..util.types import CovImputeMethod, GenoImputeMethod  # noqa: F401
from .block_hooks import BlockHookDriver, overrides
from .legacy_ops import LegacyBlockOps

_BED_MAGIC = bytes([0x6C, 0x1B, 0x01])


class Base(BlockHookDriver, LegacyBlockOps, ABC):
    #: False -> keep every block's XXz partial in HBM (one pass over the genotypes);
    #: True  -> the reference's two-pass "streaming" memory policy (recompute each block).
    _recompute_blocks = False

    def __init__(self, model: str, geno_file: str, annot_file: str = None, pheno_file: str = None,
                 cov_file: str = None, num_bin: int = 8, num_jack: int = 1, num_random_vec: int = 10,
                 geno_impute_method="binary", cov_impute_method="ignore",
                 cov_one_hot_conversion: Optional[bool] = False, categorical_threshhold: int = 100,
                 device: str = "cpu", cuda_num: Optional[int] = None, num_workers: Optional[int] = None,
                 multiprocessing: bool = True, seed: Optional[int] = None, get_trace: bool = False,
                 trace_dir: Optional[str] = None, samp_prev: Optional[float] = None,
                 pop_prev: Optional[float] = None, log: Optional[Logger] = None, kernel_path: Optional[int] = None):
        self.model = model
        self.num_jack = self.num_blocks = num_jack
        self.num_random_vec = num_random_vec
        self.num_bin = num_bin if annot_file is None else None
        self.geno_impute_methods = getattr(geno_impute_method, "value", geno_impute_method)
        self.log = log if log is not None else Logger(debug_mode=False)
        self.multiprocessing = multiprocessing
        self.kernel_path = kernel_path

        self.device_name, self.cuda_num = device, cuda_num
        self._init_device(device, cuda_num)

        # base.py:72-73 -- `seed=None` keeps the reference rule; str seeds from the CLI are coerced (Q6).  One process
        # per GPU: every rank must draw the same Z / annotation / imputation uniforms, so rank 0's seed wins.
        self.seed = self._agree_on_seed(int(time.process_time()) if seed is None else int(seed))
        np.random.seed(self.seed)
        self._check_workers(num_workers)

        self.geno_file = geno_file
        self.num_indv_original, fam_df = read_fam(geno_file + ".fam")
        self.num_snp = read_bim(geno_file + ".bim")
        self._open_bed(geno_file + ".bed")

        if annot_file is None:
            if self.num_bin is None:
                raise ValueError("Must specify number of bins if annot file is not provided")
            annot_file = "generated_annot"
            # consumes the global RNG before Z on every rank (same draws everywhere); only rank 0 writes the file
            generate_annot(annot_file if self._rank == 0 else os.devnull, self.num_snp, self.num_bin)
            self._barrier()
        self.num_bin, self.annot_matrix, self.len_bin = read_annot(annot_file, self.num_jack)

        self.pheno_file = pheno_file
        self.binary_pheno = False
        if pheno_file is not None:
            self.pheno, missing_indv, self.binary_pheno = read_pheno(pheno_file)
        else:
            self.pheno, missing_indv = None, []
        self.num_traits = self.pheno.shape[1]
        self.log._log(f"Number of traits: {self.num_traits}")

        if cov_file is None:
            self.use_cov, self.cov_matrix, self.Q = False, None, None
            self.missing_indv = missing_indv
        else:
            self.use_cov = True
            self.cov_matrix, self.missing_indv = read_cov(
                cov_file, missing_indvs=missing_indv, cov_impute_method=getattr(cov_impute_method, "value", cov_impute_method),
                one_hot_conversion=cov_one_hot_conversion, categorical_threshold=categorical_threshhold, logger=self.log)
            self.log._log(f"Rank of the covariate matrix: {np.linalg.matrix_rank(self.cov_matrix)}")
            self.Q = np.linalg.pinv(self.cov_matrix.T @ self.cov_matrix)
        if self.pheno is not None:
            self.pheno = np.delete(self.pheno, self.missing_indv, axis=0)
            self.pheno = self.pheno - np.mean(self.pheno, axis=0)

        self.num_indv = self.num_indv_original - len(self.missing_indv)
        for idx, row in enumerate(self.missing_indv, start=1):
            self.log._log(f"missing individual {idx}: FID:{fam_df.iloc[row, 0]} IID:{fam_df.iloc[row, 1]}")
        self.log._log(f"Number of individuals after filtering: {self.num_indv}")
        if self.cov_matrix is not None:
            self.log._log(f"Number of covariates: {self.cov_matrix.shape[1]}")
        self.log._log("*****")
        for i, n in enumerate(self.len_bin):
            self.log._log(f"Number of features in bin {i} : {n}")

        self.all_zb = np.random.randn(self.num_indv, self.num_random_vec)     # base.py:176
        if self.use_cov:
            self.all_Uzb = self.cov_matrix @ self.Q @ (self.cov_matrix.T @ self.all_zb)

        self.get_trace, self.trace_dir = get_trace, trace_dir
        self.samp_prev, self.pop_prev = samp_prev, pop_prev
        self.pheno_cp = self.pheno.copy()
        self.num_gen_env_bin = 0
        self.num_env = 0
        self.env = None
        self._pieces = None
        self._engine = None
        self._state = None
        self._hook_mode = False

    # ------------------------------------------------------------------ construction helpers
    def _process_group_ready(self):
        if self._world <= 1:
            return False
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=self.device)
        return True

    def _agree_on_seed(self, seed: int) -> int:
        """Multi-GPU runs are one process per rank: broadcast rank 0's seed so that every rank generates the same
        random vectors (an all-reduce over statistics of different Z would be silently wrong)."""
        if not self._process_group_ready():
            return seed
        import torch
        import torch.distributed as dist
        t = torch.tensor([seed], dtype=torch.int64, device=self.device if dist.get_backend() == "nccl" else "cpu")
        dist.broadcast(t, src=0)
        return int(t.item())

    def _barrier(self):
        if self._process_group_ready():
            import torch.distributed as dist
            dist.barrier()

    def _init_device(self, device, cuda_num):
        """base.py:195-206.  The device string is accepted for compatibility; the block kernels
        always run on CUDA (`cuda_num` or, under torchrun, LOCAL_RANK) and fail loudly otherwise."""
        import torch
        self._rank = int(os.environ.get("RANK", "0"))
        self._world = int(os.environ.get("WORLD_SIZE", "1"))
        if cuda_num is not None and cuda_num > -1:
            index = cuda_num
        else:
            index = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = torch.device("cuda", index)
        if self._world > 1 and torch.cuda.is_available():
            # one process per GPU: keep this rank's staging threads and pinned ring on its GPU's NUMA node
            from ..util.numa import bind_to_gpu_node
            self.log._debug(f"NUMA binding: {bind_to_gpu_node(index)}")
        if device == "cpu":
            self.log._debug("device='cpu' requested: pyrhe_b200 runs the block path on CUDA regardless")

    def _check_workers(self, num_workers):
        """base.py:78-96 (validation kept; workers are not used -- blocks are sharded over GPUs)."""
        import torch
        total = (torch.cuda.get_device_properties(0).multi_processor_count
                 if torch.cuda.is_available() else os.cpu_count())
        if num_workers is not None and num_workers > total:
            raise ValueError(f"The device only have {total} cores but tried to specify {num_workers} workers")
        self.num_workers = num_workers if num_workers is not None else max(total // 10, 1)
        self.log._debug(f"Number of workers: {self.num_workers}")

    def _open_bed(self, path):
        """Memory-map the SNP-major payload (replaces bed_reader.open_bed, base.py:100)."""
        row_bytes = (self.num_indv_original + 3) // 4
        with open(path, "rb") as fh:
            if fh.read(3) != _BED_MAGIC:
                raise ValueError(f"{path} is not a SNP-major PLINK 1 .bed file")
        self.geno_bed = np.memmap(path, dtype=np.uint8, mode="r", offset=3, shape=(self.num_snp, row_bytes))

    # ------------------------------------------------------------------ block interface (abstract)
    @abstractmethod
    def get_num_estimates(self):
        ...

    @abstractmethod
    def get_M_last_row(self):
        ...

    @abstractmethod
    def b_trace_calculation(self, k, j, b_idx):
        ...

    @abstractmethod
    def run(self, method):
        ...

    def pre_compute_jackknife_bin(self, j, all_gen, worker_num=None):
        """The model's per-block hook (base.py:453-455; rhe.py:13-22).  The built-in models never get here: their
        whole triple loop is fused into `rhe_block_accumulate`.  A subclass that overrides it is driven by
        `pre_compute` exactly as the reference drives it (`base/block_hooks.py`)."""
        raise NotImplementedError("built-in models run the fused block path (pyrhe_b200.engine); override this hook "
                                  "in a subclass to run your own per-block statistics")

    def _plan(self) -> PathPlan:
        """Column / estimate layout of this model for the CUDA library."""
        C = self.cov_matrix.shape[1] if self.use_cov else 0
        return PathPlan(model=self._plan_model(), K=self.num_bin, B=self.num_random_vec, C=C,
                        Ty=self.num_traits, genie_model=getattr(self, "genie_model", "G+GxE+NxE"))

    def _plan_model(self) -> str:
        return "rhe"

    # ------------------------------------------------------------------ the hot path
    def _distribute_work(self, num_jobs, num_workers):
        """base.py:530-533 -- contiguous ranges of ceil(jobs / workers)."""
        per = int(np.ceil(num_jobs / num_workers))
        return [(i * per, min((i + 1) * per, num_jobs)) for i in range(num_workers)]

    def shared_memory(self):
        """Names, shapes and dtypes of the state arrays (base.py:419-429)."""
        E, J, B, N = self.get_num_estimates(), self.num_jack, self.num_random_vec, self.num_indv
        arrays = {"XXz": ((E, J + 1, B, N), "float64"), "yXXy": ((E, J + 1), "float64"), "M": ((J + 1, E), "int64")}
        if self.use_cov:
            arrays.update({"UXXz": ((E, J + 1, B, N), "float64"), "XXUz": ((E, J + 1, B, N), "float64")})
        self.shared_memory_arrays = arrays
        return arrays

    def _setup_shared_memory(self):
        """base.py:439-450 (name kept): only the small M table lives on the host now."""
        self.num_estimates = self.get_num_estimates()
        self.M = np.zeros((self.num_jack + 1, self.num_estimates), dtype=np.int64)
        self.M[self.num_jack] = self.get_M_last_row()

    def pre_compute(self):
        """One pass over the packed genotypes on the GPU(s); replaces base.py:535-554 + aggregate.

        The Z-dependent statistics do not depend on the trait, so all phenotype columns ride
        through the same pass and later traits reuse the pieces."""
        self._setup_shared_memory()
        self._state = None
        self._hook_mode = self._uses_block_hooks()
        if self._hook_mode:                       # an extender's own block hook: the reference's loop around it
            if self._world > 1:
                raise NotImplementedError("custom block hooks run in single-process mode (the built-in models shard "
                                          "their fused path over GPUs)")
            t0 = time.time()
            if self._recompute_blocks:
                self._hook_pre_compute_streaming()
            else:
                self._hook_pre_compute()
            self.log._debug(f"Precompute total time: {time.time() - t0}")
            return
        if self._pieces is not None:
            self.M = self._pieces["M"].copy()
            return
        from ..engine import RheEngine
        t0 = time.time()
        plan = self._plan()
        keep = np.ones(self.num_indv_original, dtype=bool)
        keep[list(self.missing_indv)] = False
        self._host_terms, Y_res = host_terms(plan, self.all_zb, self.cov_matrix, self.pheno_cp, self._env_vector())
        self._Y_res = Y_res
        pg = None
        self._process_group_ready()
        # int8 tcgen05 kernels (RHE, RHE-DOM with its [g == 2] operand, GENIE with the env-scaled GxE set) whenever the
        # layout fits them; otherwise the engine falls back to the CUDA-core kernels of the same library with a warning
        path = self.kernel_path
        if path is None and "PYRHE_B200_PATH" in os.environ:
            path = int(os.environ["PYRHE_B200_PATH"])
        eng = RheEngine(plan, n_indv=self.num_indv_original, keep=keep, annot=self.annot_matrix,
                        num_jack=self.num_jack, impute=self.geno_impute_methods, seed=self.seed, device=self.device,
                        kernel_path=path, rank=self._rank, world=self._world,
                        store_partials=not self._recompute_blocks, process_group=pg,
                        retile=os.environ.get("PYRHE_B200_RETILE", "1") != "0")
        eng.set_rhs(self.all_zb, self.cov_matrix, Y_res, self._env_vector())
        # bounded-memory ingest overlapped with compute: block j+1 is staged and copied while
        # block j runs; the rank's `.bed` share stays resident only when it fits the HBM, otherwise it streams
        # through a ring of block slots (and, for the streaming policy, streams a second time)
        ring = os.environ.get("PYRHE_B200_RING_BLOCKS")       # force the bounded ring (default: only when it must)
        streamer = eng.stream_genotypes(self.geno_bed, ring_blocks=int(ring) if ring else "auto")
        # the host half of the normal equations that needs only the per-bin Gram pieces runs while the device still
        # forms the leave-one-out Grams (trait-independent; `_solve_all` finishes it per trait)
        ht = self._host_terms
        try:
            self._pieces = eng.run(upload=streamer,
                                   gram_hook=lambda G: normal_equations_prepare(plan, ht, loo_grams(G), eng.Mjk))
        finally:
            streamer.close()
        self.ingest_report = dict(ring_blocks=eng.ring_blocks, genotype_bytes=eng.genotype_bytes(),
                                  staged_bytes=streamer.bytes_staged, passes=streamer.passes)
        self._engine = eng
        self._plan_cached = plan
        self._G_tot = self._pieces["G_blk"].sum(axis=0)
        self.M = self._pieces["M"].copy()
        assert np.array_equal(self.M[self.num_jack], np.asarray(self.get_M_last_row()))
        self.log._debug(f"Precompute total time: {time.time() - t0}")
        self.aggregate()

    def aggregate(self):
        """base.py:465-500 -- totals and leave-one-out sums are formed on the device
        (all-reduce of S, `S - P_j` inside `rhe_loo_gram`); nothing is left to do on the host."""

    def _env_vector(self):
        return None if self.env is None else np.asarray(self.env).reshape(-1)

    def _trait_index(self) -> int:
        return getattr(self, "_trait", 0)

    def _uses_block_hooks(self) -> bool:
        """Has a subclass supplied its own per-block hook(s) (base.py:453-455, base_streaming.py:106-108)?"""
        from .base_streaming import StreamingBase
        own1 = overrides(self, "pre_compute_jackknife_bin", Base, StreamingBase)
        if not self._recompute_blocks:
            return own1
        own2 = overrides(self, "pre_compute_jackknife_bin_pass_2", Base, StreamingBase)
        if own1 != own2:
            raise TypeError("a streaming model must override both pre_compute_jackknife_bin(j, all_gen, worker_num) "
                            "and pre_compute_jackknife_bin_pass_2(j, all_gen) (base_streaming.py:85-144)")
        return own1

    def setup_lhs_rhs_jackknife(self, j, trace_sums, is_streaming=False):
        """(T, q) of jackknife sample j (j == num_jack: all SNPs) -- base.py:568-628."""
        if self._hook_mode:
            return self._hook_lhs_rhs(j, trace_sums, is_streaming)
        plan, pc = self._plan_cached, self._pieces
        G_loo = self._G_tot - pc["G_blk"][j] if j < self.num_jack else self._G_tot
        T, q = normal_equations(plan, self._host_terms, pc["XX"][j], G_loo, self.M[j], trait=self._trait_index())
        if self.get_trace and trace_sums is not None:
            trace_sums[j] = trace_sums_row(T, self.M[j], self.num_indv, self.num_estimates)
        return T, q

    def solve_linear_equation(self, X, y):
        return stats.solve_lstsq(X, y)

    def solve_linear_qr(self, X, y):
        return stats.solve_qr(X, y)

    def _solve_all(self, method):
        """Shared body of `estimate` (base.py:645-671): returns (sigma [J+1, E+1], T column of traces)."""
        trace_sums = (np.zeros((self.num_jack + 1, self.num_estimates, self.num_estimates))
                      if self.get_trace else None)
        sigmas, trace_cols = [], []
        batched = type(self).setup_lhs_rhs_jackknife is Base.setup_lhs_rhs_jackknife and not self._hook_mode
        if batched:       # all J + 1 systems assembled in one vectorised pass (an extender's override is honoured below)
            pc = self._pieces
            if pc.get("gram_hook") is not None and np.array_equal(self.M, pc["M"]):
                T_all, q_all = normal_equations_finish(pc["gram_hook"], pc["XX"], trait=self._trait_index())
            else:
                T_all, q_all = normal_equations_batch(self._plan_cached, self._host_terms, pc["XX"],
                                                      loo_grams(pc["G_blk"]), self.M, trait=self._trait_index())
        for j in range(self.num_jack + 1):
            jj = 1 if (self.num_jack == 1 and j == 0) else j          # base.py:654-655
            if batched:
                T, q = T_all[jj], q_all[jj]
                if trace_sums is not None:
                    trace_sums[jj] = trace_sums_row(T, self.M[jj], self.num_indv, self.num_estimates)
            else:
                T, q = self.setup_lhs_rhs_jackknife(jj, trace_sums)
            sigmas.append(stats.solve(T, q, method))
            trace_cols.append(T[:, self.num_estimates].copy())
        if self.get_trace:
            self.get_trace_summary(trace_sums)
        return np.array(sigmas), np.array(trace_cols)

    def estimate(self, method: str = "lstsq") -> Tuple[List[List], List]:
        if self._hook_mode and self._recompute_blocks:
            return self._hook_estimate_streaming(method)
        sigma, _ = self._solve_all(method)
        return sigma[:-1, :], sigma[-1, :]

    def estimate_error(self, ests):
        return stats.jackknife_se(ests, self.num_jack)

    def compute_h2_nonoverlapping(self, sigma_est_jackknife, sigma_ests_total):
        h2 = stats.h2_nonoverlapping(np.vstack([sigma_est_jackknife, sigma_ests_total[np.newaxis, :]]),
                                     self.num_estimates)
        return h2[:-1, :], h2[-1, :]

    def compute_h2_overlapping(self, sigma_est_jackknife, sigma_ests_total):
        if not hasattr(self, "_cooc"):
            self._cooc = stats.bin_cooccurrence(self.annot_matrix, block_ranges(self.num_snp, self.num_jack))
        h2 = stats.h2_overlapping(np.vstack([sigma_est_jackknife, sigma_ests_total[np.newaxis, :]]), self.M,
                                  self._cooc[0], self._cooc[1], self.num_estimates)
        return h2[:-1, :], h2[-1, :]

    def compute_enrichment(self, h2_jackknife, h2_total):
        en = stats.enrichment(np.vstack([h2_jackknife, h2_total[np.newaxis, :]]), self.M, self.num_estimates)
        return en[:-1, :], en[-1, :]

    @staticmethod
    def _calc_lsum(tr, n, m1, m2):
        return (tr - n) * (m1 * m2) / pow(n, 2)

    def get_trace_summary(self, trace_sums):
        prefix = stats.write_trace_files(trace_sums, self.M, pheno_file=self.pheno_file, trace_dir=self.trace_dir,
                                         num_indv=self.num_indv, num_snp=self.num_snp, num_jack=self.num_jack,
                                         num_bin=self.num_bin, num_random_vec=self.num_random_vec)
        self.log._log(f"Saved trace summary into {prefix}(.tr/.MN)")
        self.log._debug(f"Trace saved to {prefix}.tr")
        self.log._debug(f"MN data saved to {prefix}.MN")

    def _compute_liability_h2(self, h2, seh2):
        return stats.liability_h2(h2, seh2, self.samp_prev, self.pop_prev)

    # the reference's run() calls this (undefined) name -- SURVEY.md §9.3 Q10
    calculate_liability_h2 = _compute_liability_h2

    def regress_pheno(self, cov_matrix, pheno):
        """base.py:396-401."""
        Q = np.linalg.pinv(cov_matrix.T @ cov_matrix)
        return pheno - cov_matrix @ (Q @ (cov_matrix.T @ pheno))

    def _finalize(self):
        """base.py:556-563 (shared-memory teardown in the reference): releases the GPU context."""
        if self._engine is not None:
            self._engine.close()
            self._engine = None
        if getattr(self, "_dec_engine", None) is not None:
            self._dec_engine.close()
            self._dec_engine = None

    def __call__(self, trait, method: str = "QR"):
        self._trait = trait
        self.pheno = self.pheno_cp[:, trait].reshape(-1, 1)
        self.log._log("*****")
        self.log._log(f"OUTPUT FOR TRAIT {trait}: ")
        self.pre_compute()
        res = self.run(method=method)
        if trait >= self.num_traits - 1:
            self._finalize()
        return res

    # ------------------------------------------------------------------ shared report helpers
    def _log_variance_components(self, sigma_total, sig_errs):
        self.log._log("Variance components: ")
        last = len(sigma_total) - 1
        for i, est in enumerate(sigma_total):
            label = "Sigma^2_e" if i == last else f"Sigma^2_g[{i}]"
            self.log._log(f"{label} : {est}  SE : {sig_errs[i]}")

    def _log_h2_block(self, h2_total, h2_errs):
        last = len(h2_total) - 1
        for i, est in enumerate(h2_total):
            if i == last:
                self.log._log(f"Total h2 : {est} SE: {h2_errs[i]}")
            else:
                self.log._log(f"h2_g[{i}] : {est} : {h2_errs[i]}")

    def _log_enrichment_block(self, enr_total, enr_errs):
        for i, est in enumerate(enr_total):
            self.log._log(f"Enrichment g[{i}] : {est} SE : {enr_errs[i]}")
