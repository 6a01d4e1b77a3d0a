"""The reference's *block interface* for extenders, and its state arrays.

/root/reference/pyrhe/src/base/base.py:453-463,503-527 and base_streaming.py:85-144 define how a model plugs into the
framework: the worker decodes one jackknife block, imputes it, splits it by bin and hands the list `all_gen` to the
model's `pre_compute_jackknife_bin(j, all_gen)` hook, which fills the state arrays `XXz, UXXz, XXUz, yXXy, M`
(shapes base.py:419-429 / docs/api/base.rst:75-116); `aggregate` turns them into leave-one-out sums and
`setup_lhs_rhs_jackknife` reads them back.

The built-in models never go through the hooks (their whole triple loop is one fused kernel sequence per block), but
the interface is kept for two users:

* a subclass that OVERRIDES `pre_compute_jackknife_bin` (or, for a streaming model, `..._pass_2`) is detected by
  `Base.pre_compute` and driven exactly as the reference drives it: blocks decoded on the GPU, the hook called per
  block, `aggregate`, then `(T, q)` from the arrays the hook filled;
* `model.XXz / UXXz / XXUz / yXXy` on a built-in model are materialised on demand from the device state
  (`S`, `P_j`), in the reference's layout and meaning (after `aggregate`: slot j = leave-one-out sum, slot J = total).

Everything here is host-side fp64 numpy over arrays the reference itself would hold; the O(N m) work stays in
`read_geno` (GPU decode) and the `_compute_*` helpers (`mat_mul` on CUDA).
"""
from __future__ import annotations

import numpy as np

from ..hostmath import block_ranges

STATE_NAMES = ("XXz", "UXXz", "XXUz", "yXXy")


def overrides(obj, name: str, *bases) -> bool:
    """True when type(obj) resolves `name` to something other than the placeholder of the framework classes."""
    impl = getattr(type(obj), name, None)
    return impl is not None and all(impl is not getattr(b, name, None) for b in bases)


class BlockHookDriver:
    """Mixed into `Base`: the reference's per-block loop around a user-supplied hook."""

    # ------------------------------------------------------------------ state arrays (base.py:419-450)
    def _hook_alloc_state(self, streaming: bool):
        E, J, B, N = self.num_estimates, self.num_jack, self.num_random_vec, self.num_indv
        slots = 2 if streaming else J + 1          # streaming: slot 0 = total, slot 1 = current leave-one-out
        state = {"XXz": np.zeros((E, slots, B, N)), "yXXy": np.zeros((E, slots))}
        if self.use_cov:
            state["UXXz"] = np.zeros((E, slots, B, N))
            state["XXUz"] = np.zeros((E, slots, B, N))
        self._state = state

    def _hook_block(self, j):
        """decode -> impute -> bin gather of block j, as _pre_compute_worker does (base.py:510-519)."""
        subsample, sub_annot = self._get_jacknife_subsample(j)
        subsample = self.impute_geno(subsample)
        assert subsample.shape[0] == self.num_indv
        return self.partition_bins(subsample, sub_annot)

    # ------------------------------------------------------------------ non-streaming (base.py:503-554, 465-500)
    def _hook_pre_compute(self):
        self._hook_alloc_state(streaming=False)
        for j in range(self.num_jack):
            np.random.seed(self.seed)              # base.py:510: the imputation draws restart at every block
            self.pre_compute_jackknife_bin(j, self._hook_block(j))
        self._hook_aggregate()

    def _nxe_rows(self):
        """Estimate rows of the heteroscedastic-noise component (X = diag(env), base.py:472-481)."""
        n_env = getattr(self, "num_env", 0)
        has = n_env and getattr(self, "genie_model", "") == "G+GxE+NxE"
        return range(self.num_estimates - n_env, self.num_estimates) if has else range(0)

    def _hook_fill_nxe(self, slot):
        """The reference multiplies by an N x N matrix; with X diagonal, XXz = env^2 * z (same fp32 products)."""
        B = self.num_random_vec
        for k in self._nxe_rows():
            e = k - self.num_bin - self.num_gen_env_bin
            env32 = self.env[:, e].astype(np.float32)
            for b in range(B):
                self._state["XXz"][k][slot][b] = env32 * (env32 * self.all_zb[:, b].astype(np.float32))
            y = self.pheno if not self.use_cov else self.regress_pheno(self.cov_matrix, self.pheno)
            v = env32 * y[:, 0].astype(np.float32)
            self._state["yXXy"][k][slot] = np.float32(v @ v)
            if self.use_cov:                       # base.py:479-481 sits outside the loop over b: only b = B - 1
                b = B - 1
                self._state["UXXz"][k][slot][b] = self._compute_UXXz(self._state["XXz"][k][slot][b])
                self._state["XXUz"][k][slot][b] = env32 * (env32 * self.all_Uzb[:, b].astype(np.float32))

    def _hook_aggregate(self):
        """Totals into slot J, then slot j <- total - block j (base.py:465-500)."""
        J = self.num_jack
        nxe = set(self._nxe_rows())
        regular = [k for k in range(self.num_estimates) if k not in nxe]
        for name, A in self._state.items():
            A[regular, J] = A[regular, :J].sum(axis=1)
        self._hook_fill_nxe(J)
        for A in self._state.values():
            A[:, :J] = A[:, J:J + 1] - A[:, :J]

    # ------------------------------------------------------------------ streaming (base_streaming.py:85-144)
    def _hook_pre_compute_streaming(self):
        """Pass 1: the hook accumulates every block into slot 0 (`worker_num = 0`, one worker per GPU process)."""
        self._hook_alloc_state(streaming=True)
        for j in range(self.num_jack):
            np.random.seed(self.seed)              # parity target = the non-streaming rule
            self.pre_compute_jackknife_bin(j, self._hook_block(j), 0)
        self._hook_fill_nxe(0)
        for k in self._nxe_rows():
            for A in self._state.values():
                A[k][1] = A[k][0]

    def _hook_estimate_streaming(self, method):
        """Pass 2: re-decode block j, let the hook form the leave-one-out sums in slot 1, build (T, q), solve."""
        from .. import stats
        from ..assemble import trace_sums_row
        trace_sums = (np.zeros((self.num_jack + 1, self.num_estimates, self.num_estimates)) if self.get_trace else None)
        sigmas = []
        all_gen = None
        for j in range(self.num_jack + 1):
            if j != self.num_jack:
                np.random.seed(self.seed)
                all_gen = self._hook_block(j)
            self.pre_compute_jackknife_bin_pass_2(j, all_gen)
            T, q = self.setup_lhs_rhs_jackknife(j, None, is_streaming=True)
            if trace_sums is not None:
                trace_sums[j] = trace_sums_row(T, self.M[j], self.num_indv, self.num_estimates)
            sigmas.append(stats.solve(T, q, method))
        if self.get_trace:
            self.get_trace_summary(trace_sums)
        sig = np.array(sigmas)
        return sig[:-1, :], sig[-1, :]

    # ------------------------------------------------------------------ (T, q) from the arrays (base.py:568-628)
    def _hook_lhs_rhs(self, j, trace_sums, is_streaming=False):
        E, B, N = self.num_estimates, self.num_random_vec, self.num_indv
        st = self._state
        s = 1 if is_streaming else j
        XXz = st["XXz"][:, s].reshape(E, B * N)
        Mrow = self.M[j].astype(np.float64)
        V = XXz @ XXz.T
        if self.use_cov:
            W, Q = self.cov_matrix, self.Q
            proj = (st["XXz"][:, s] @ W)                                 # [E, B, C] = (W^T XXz)^T
            UX = np.einsum("ebc,cd,nd->ebn", proj, Q, W).reshape(E, B * N)   # U XXz
            r1 = UX @ XXz.T                                              # <U XXz_a, XXz_c>
            r2 = st["XXUz"][:, s].reshape(E, B * N) @ st["UXXz"][:, s].reshape(E, B * N).T
            V = V + r2 - 2 * r1
        V = V / B
        MM = np.outer(Mrow, Mrow)
        T = np.zeros((E + 1, E + 1))
        np.divide(V, MM, out=T[:E, :E], where=MM != 0)
        if self.get_trace and trace_sums is not None:
            from ..assemble import trace_sums_row
            trace_sums[j] = trace_sums_row(T, self.M[j], N, E)
        q = np.zeros((E + 1, 1))
        for k in range(E):
            tr = self.b_trace_calculation(k, j, s)
            if self.use_cov:
                tr = tr - np.sum(st["XXz"][k][s] * self.all_Uzb.T) / (B * Mrow[k])
            T[k, E] = T[E, k] = tr
            q[k] = st["yXXy"][k][s] / Mrow[k] if Mrow[k] != 0 else 0
        T[E, E] = N if not self.use_cov else N - self.cov_matrix.shape[1]
        y = self.pheno if not self.use_cov else self.regress_pheno(self.cov_matrix, self.pheno)
        q[E] = y.T @ y
        return T, q

    # ------------------------------------------------------------------ built-in models: arrays on demand
    def _materialise_state(self):
        """`XXz / UXXz / XXUz / yXXy` of a built-in model in the reference's post-`aggregate` meaning
        (base.py:419-429, 465-500): slot j < J = leave-one-out sum, slot J = total over all blocks.

        `XXz` comes from the device state of the fused path (totals S and stored block partials P_j; a streaming
        model re-runs the pass with stored partials); `UXXz = W Q W^T XXz` is formed from it; `XXUz = X X^T (U z)` is
        a second pass of the same kernels with `U Z` as the random vectors (the fused path itself never needs these
        N-vectors); `yXXy` is the (y, y) entry of the per-block Grams."""
        if self._world > 1:
            raise NotImplementedError("state arrays are materialised in single-process runs only")
        import torch
        from ..engine import RheEngine
        from ..hostmath import host_terms
        plan, J, N = self._plan(), self.num_jack, self.num_indv
        keep = np.ones(self.num_indv_original, dtype=bool)
        keep[list(self.missing_indv)] = False
        env = self._env_vector()

        def vectors(Z):
            ht, Y_res = host_terms(plan, Z, self.cov_matrix, self.pheno_cp, env)
            eng = RheEngine(plan, n_indv=self.num_indv_original, keep=keep, annot=self.annot_matrix, num_jack=J,
                            impute=self.geno_impute_methods, seed=self.seed, device=self.device,
                            kernel_path=self.kernel_path, store_partials=True)
            try:
                eng.set_rhs(Z, self.cov_matrix, Y_res, env)
                eng.load_genotypes(self.geno_bed)
                pieces = eng.run()
                S = eng.S.cpu().numpy()[..., : self.num_indv_original][..., keep].astype(np.float64)
                P = eng.P_all.cpu().numpy()[..., : self.num_indv_original][..., keep].astype(np.float64)
            finally:
                eng.close()
            torch.cuda.empty_cache()
            out = np.concatenate([S[None] - P, S[None]], axis=0).transpose(1, 0, 2, 3)   # [E, J + 1, B, N]
            return np.ascontiguousarray(out), pieces

        XXz, pieces = vectors(self.all_zb)
        state = {"XXz": XXz}
        yc = plan.col_y(self._trait_index())
        G = pieces["G_blk"]
        yy = np.zeros((plan.E, J + 1))
        yy[: plan.E_reg, J] = G[:, :, yc, yc].sum(axis=0)
        yy[: plan.E_reg, :J] = yy[: plan.E_reg, J:J + 1] - G[:, :, yc, yc].T
        if plan.has_nxe:
            y_res = self._host_terms_for_state(plan, env)
            yy[plan.E_reg, :] = y_res
        state["yXXy"] = yy
        if self.use_cov:
            W, Q = self.cov_matrix, self.Q
            state["UXXz"] = np.einsum("ejbn,nc,cd,md->ejbm", XXz, W, Q, W, optimize=True)
            XXUz, _ = vectors(self.all_Uzb)
            if plan.has_nxe:                       # base.py:479-481: only b = B - 1 of the NxE row is filled
                state["UXXz"][plan.E_reg, :, :-1] = 0
                XXUz[plan.E_reg, :, :-1] = 0
            state["XXUz"] = XXUz
        self._state = state

    def _host_terms_for_state(self, plan, env):
        from ..hostmath import host_terms
        ht, _ = host_terms(plan, self.all_zb, self.cov_matrix, self.pheno_cp, env)
        return ht.nxe_yxxy[self._trait_index()]

    def _state_array(self, name):
        st = getattr(self, "_state", None)
        if st is None or name not in st:
            if name in ("UXXz", "XXUz") and not self.use_cov:
                raise AttributeError(f"{name} exists only with covariates (base.py:424-428)")
            if getattr(self, "_pieces", None) is None and st is None:
                raise AttributeError(f"{name} is allocated by pre_compute() (base.py:439-450)")
            if st is None or name not in st:
                self._materialise_state()
        return self._state[name]


def _state_property(name):
    def get(self):
        return self._state_array(name)

    def set_(self, value):                         # the streaming reference rebinds the arrays (base_streaming.py:34-37)
        if getattr(self, "_state", None) is None:
            self._state = {}
        self._state[name] = value

    return property(get, set_, doc=f"reference state array `{name}` (base.py:419-429)")


for _n in STATE_NAMES:
    setattr(BlockHookDriver, _n, _state_property(_n))
