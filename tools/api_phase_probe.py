"""Where the seconds of a model-API call go (`StreamingRHE(...)(trait=0)` on a config-2 `.bed` in tmpfs): engine set-up,
genotype allocation, staging, the pass itself, the estimator tail.  Wraps the engine's methods with timers."""
import os
import shutil
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from pyrhe_b200 import engine as E

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "config2"])
T = {}


def timed(cls, name, sync=False):
    fn = getattr(cls, name)

    def wrap(*a, **k):
        t0 = time.perf_counter()
        out = fn(*a, **k)
        if sync:
            torch.cuda.synchronize(dev)
        T[name] = T.get(name, 0.0) + time.perf_counter() - t0
        return out
    setattr(cls, name, wrap)


for name in ("__init__", "set_rhs", "alloc_genotypes", "stream_genotypes", "run", "reserve_state"):
    timed(E.RheEngine, name, sync=name in ("alloc_genotypes", "set_rhs"))
timed(E.RheEngine, "close")
timed(E.BlockStreamer, "_read_chunk")
timed(E.BlockStreamer, "__init__")
timed(E.BlockStreamer, "acquire")

outdir = tempfile.mkdtemp(prefix="pyrhe_probe_", dir="/dev/shm")
try:
    paths = bench.write_synthetic_plink(outdir, wl, dev)
    import pyrhe.models as models
    from pyrhe.src.util import Logger
    for cls_name, ring in (("StreamingRHE", None), ("RHE", None), ("RHE", 4), ("StreamingRHE", None)):
        if ring is None:
            os.environ.pop("PYRHE_B200_RING_BLOCKS", None)
        else:
            os.environ["PYRHE_B200_RING_BLOCKS"] = str(ring)
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        T.clear()
        model = getattr(models, cls_name)(model="rhe", num_jack=wl["J"], num_random_vec=wl["B"], seed=0,
                                          geno_impute_method="binary", device="cuda", num_workers=1,
                                          log=Logger(suppress=True, debug_mode=False), **paths)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        t_pc = time.perf_counter()
        model.pre_compute()
        torch.cuda.synchronize(dev)
        t_pc = time.perf_counter() - t_pc
        t1 = time.perf_counter()
        res = model.run(method="QR")
        t2 = time.perf_counter()
        model._finalize()
        t3 = time.perf_counter()
        T["estimator tail (run)"] = t2 - t1
        T["_finalize"] = t3 - t2
        t_call = time.perf_counter() - t0
        print(f"{cls_name} ring={ring}: call {t_call:.3f} s (pre_compute {t_pc:.3f}); "
              + ", ".join(f"{k} {v:.3f}" for k, v in T.items()) + "  (_read_chunk: summed over the staging threads)")
        del model
finally:
    shutil.rmtree(outdir, ignore_errors=True)
