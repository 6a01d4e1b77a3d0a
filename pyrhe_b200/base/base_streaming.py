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
        """Reference hook (streaming_rhe.py:28-43); fused into the engine's recompute pass."""
        raise NotImplementedError("built-in models run the fused block path (pyrhe_b200.engine)")
