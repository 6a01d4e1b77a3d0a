"""`StreamingBase`: the reference's memory-lean two-pass variant
(/root/reference/pyrhe/src/base/base_streaming.py).

API and MRO are kept (`StreamingRHE(RHE, StreamingBase)` ...).  The mechanism becomes a
memory *policy* of the device engine: pass 1 accumulates only the totals S, pass 2 recomputes
each block's partial P_j and forms the leave-one-out Gram at once, so no per-block partial is
ever stored (base_streaming.py:85-144 decodes and multiplies every block twice as well).

Parity target: the streaming classes reproduce the NON-streaming
reference, i.e. the mathematically intended value.  The reference's streaming code accumulates
`U @ (running sum)` (streaming_rhe.py:21), which makes its covariate terms depend on the worker
count, and only runs with multiprocessing and covariates; this implementation works in every mode.
"""
from .base import Base


class StreamingBase(Base):
    _recompute_blocks = True

    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    def pre_compute_jackknife_bin_pass_2(self, j, all_gen):
        """Pass-2 hook (base_streaming.py:106-108; streaming_rhe.py:28-43): leave-one-out sums of jackknife sample j into
        slot 1 of the state arrays (slot 0 holds the totals of pass 1).  The built-in models fuse it into the engine's
        recompute sweep; a subclass that overrides it (together with the three-argument
        `pre_compute_jackknife_bin(j, all_gen, worker_num)`) is driven as in the reference (`base/block_hooks.py`)."""
        raise NotImplementedError("built-in models run the fused block path (pyrhe_b200.engine)")

    def shared_memory(self):
        """base_streaming.py:61-83, for one worker per GPU process: the state an extender's hooks see is
        `XXz, UXXz, XXUz [E, 2, B, N]` and `yXXy [E, 2]` -- slot 0 the totals (pass 1 accumulates with
        `worker_num = 0`), slot 1 the current leave-one-out sums (pass 2)."""
        E, B, N = self.num_estimates, self.num_random_vec, self.num_indv
        arrays = {"XXz": ((E, 2, B, N), "float64"), "yXXy": ((E, 2), "float64"),
                  "M": ((self.num_jack + 1, E), "int64")}
        if self.use_cov:
            arrays.update({"UXXz": ((E, 2, B, N), "float64"), "XXUz": ((E, 2, B, N), "float64")})
        self.shared_memory_arrays = arrays
        return arrays
