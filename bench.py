"""Benchmark of the RHE trace-estimation hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 3 --warmup 3                       # this implementation
    python -m torch.distributed.run --nproc-per-node 8 ... bench.py --gpus 8 ...
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1      # CPU arm (the unmodified reference from oracle/_ref)

A "step" is one full RHE jackknife over synthetic genotypes of the named shape: every jackknife
block of every rank through the block kernels, the all-reduce of the totals, the leave-one-out
Grams, the D2H of the Gram pieces, host assembly of the J+1 normal equations and their solves.
`value` = packed .bed bytes (ceil(N0/4) * M) / step time with the ingested genotypes resident in HBM
(packed rows + the per-SNP allele counts the ingest kernel takes when a block lands);
`e2e` = the same job through the engine's ingest API (`stream_genotypes`: host rows -> staging threads ->
pinned ring -> PCIe -> device, counts on arrival) with every byte crossing host memory and PCIe inside the
timed region (with several GPUs whose links differ by more than 10 % -- `h2d_ceiling.per_rank_gbs`, measured with all
ranks copying at once -- the leg runs a second time with block shares proportional to the measured rates and both legs
are recorded); `e2e_model_api` = `StreamingRHE(...)(trait=0)` on a real `.bed` file at config-2 size (first call of the
process = `value`, the same call again = `warm_value`);
`first_pass_ms` = the very first pass of a fresh context (what a model run actually executes), and
`step_recount_ms` the step when every block's allele counts are re-taken inside it (three reads per block,
the round-1 structure).  The problem size is fixed as GPUs are added ("scaling": "strong"), as
BASELINE.json quotes the metric for one problem at 1/2/4/8 GPUs.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[4]: the configuration the metric is quoted on
    "config5": dict(N=500_000, M=1_000_000, J=100, K=8, C=5, B=10, model="rhe"),
    "config2": dict(N=200_000, M=500_000, J=100, K=8, C=5, B=10, model="rhe"),
    # BASELINE.json configs[2], configs[3] (parity-scale runs of the other two model families)
    "config3": dict(N=200_000, M=500_000, J=100, K=8, C=5, B=10, model="rhe_dom"),
    "config4": dict(N=300_000, M=500_000, J=100, K=8, C=5, B=10, model="genie"),
    # a few config2 / config5 sized blocks: short enough to run under ncu
    "profile": dict(N=200_000, M=20_000, J=4, K=8, C=5, B=10, model="rhe"),
    "profile5": dict(N=500_000, M=40_000, J=4, K=8, C=5, B=10, model="rhe"),
    "profile3": dict(N=200_000, M=20_000, J=4, K=8, C=5, B=10, model="rhe_dom"),
    "profile4": dict(N=300_000, M=20_000, J=4, K=8, C=5, B=10, model="genie"),
    "small": dict(N=20_000, M=40_000, J=20, K=8, C=5, B=10, model="rhe"),
}
METRIC = "rhe_genotype_throughput"
UNIT = "GB/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": "warm-up + timed steps"}



# ----------------------------------------------------------------------------------------------
def cpu_arm(wl, args):
    """The CPU side of the comparison on the host cores: the UNMODIFIED reference from oracle/_ref (kind "reference") when
    the shipped copy is intact, else the oracle port (kind "port").  Returns (baseline object, kind)."""
    from oracle import ref_baseline
    why = ref_baseline.available()
    if not why and not args.cpu_port:
        return ref_baseline.RefBaseline(wl["N"], wl["K"], wl["C"], wl["B"], snps_per_block=args.ref_snps), "reference"
    from oracle.cpu_baseline import CpuBaseline
    return CpuBaseline(wl["N"], wl["K"], wl["C"], wl["B"], snps_per_block=args.cpu_snps), "port"


def reference_arm(args, wl, rank):
    """CPU implementation of the path -- the reference's own code when oracle/_ref is present --, all host cores, bounded
    sample per step (each step of the reference takes about half a minute: at most one warm-up step is run)."""
    if rank != 0:
        return
    def measure():
        cb, kind = cpu_arm(wl, args)
        try:
            for _ in range(min(args.warmup, 1)):
                cb.step()
            secs, geno = 0.0, 0.0
            for _ in range(args.steps):
                s, _, g = cb.step()
                secs += s
                geno += g
        finally:
            cb.close()
        return cb, kind, secs, geno

    try:
        cb, kind, secs, geno = measure()
    except Exception as exc:                                 # e.g. the host cannot hold the reference's state arrays
        if args.cpu_port:
            raise
        sys.stderr.write(f"reference arm: the unmodified reference failed ({exc}); timing the oracle port instead\n")
        args.cpu_port = True
        cb, kind, secs, geno = measure()
    gbs = geno / 4 / secs / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, **wl, "sample": cb.describe()},
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": cb.cores, "kind": kind, "sample": cb.describe()},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "projected_full_job_s": wl["N"] / 4 * wl["M"] / 1e9 / gbs,
    }
    _emit(line)


# ----------------------------------------------------------------------------------------------
def write_synthetic_plink(outdir, wl, dev, chunk_snps=4096):
    """A real PLINK data set of the workload's shape on local storage: `.bed` rows generated on the device
    (`rhe_synth_genotypes`, the generator of the resident leg) and written out, text files as the reference reads them."""
    import torch
    from pyrhe_b200 import _lib, synth
    lib = _lib.load()
    N, M, K, Cc = wl["N"], wl["M"], wl["K"], wl["C"]
    rb = (N + 3) // 4
    pitch = (rb + 127) // 128 * 128
    prefix = os.path.join(outdir, "bench")
    st = torch.cuda.current_stream(dev)
    buf = torch.zeros((chunk_snps, pitch), dtype=torch.uint8, device=dev)
    host = torch.empty((chunk_snps, rb), dtype=torch.uint8).pin_memory()
    with open(prefix + ".bed", "wb") as f:
        f.write(synth.BED_MAGIC)
        for s0 in range(0, M, chunk_snps):
            m = min(chunk_snps, M - s0)
            _lib.check(lib.rhe_synth_genotypes(C.c_void_p(buf.data_ptr()), m, pitch, N, s0, 1234, 0.001,
                                               C.c_void_p(st.cuda_stream)))
            host[:m].copy_(buf[:m, :rb])
            f.write(host[:m].numpy().tobytes())
    rng = np.random.default_rng(1)
    ids = np.arange(N)
    np.savetxt(prefix + ".fam", np.stack([ids, ids, 0 * ids, 0 * ids, 0 * ids, 0 * ids - 9], 1), fmt="%d")
    with open(prefix + ".bim", "w") as f:
        f.write("".join(f"1\trs{i}\t0\t{i}\tA\tG\n" for i in range(M)))
    annot = synth.random_annot(M, K, rng)
    np.savetxt(prefix + ".annot", annot, fmt="%d")
    y = rng.standard_normal(N)
    with open(prefix + ".pheno", "w") as f:
        f.write("FID IID pheno0\n")
        f.write("".join(f"{i} {i} {v!r}\n" for i, v in enumerate(y.tolist())))
    W = rng.standard_normal((N, Cc))
    W[:, 0] = rng.random(N) < 0.5
    with open(prefix + ".cov", "w") as f:
        f.write("FID IID " + " ".join(f"cov{c}" for c in range(Cc)) + "\n")
        np.savetxt(f, np.concatenate([ids[:, None], ids[:, None], W], 1), fmt=["%d", "%d"] + ["%.17g"] * Cc)
    return dict(geno_file=prefix, annot_file=prefix + ".annot", pheno_file=prefix + ".pheno", cov_file=prefix + ".cov")


def model_api_e2e(args, dev):
    """`e2e_model_api`: the call a user makes -- `StreamingRHE(**files)(trait=0)` -- on a real `.bed` file at the size
    of BASELINE.json's single-GPU configuration, read from local storage inside the timed region (file -> staging
    threads -> pinned ring -> PCIe -> HBM -> kernels -> estimates).  Run twice more with the genotype ring forced
    (`PYRHE_B200_RING_BLOCKS=4`): HBM use bounded by four block slots whatever the size of the file, once with stored
    partials (one pass over the file) and once with the streaming policy (two passes)."""
    import gc
    import shutil
    import tempfile
    import torch
    wl = dict(WORKLOADS[args.api_workload])
    bed_bytes = float((wl["N"] + 3) // 4) * wl["M"]
    root = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 1.5 * bed_bytes + 4e9 else None
    if root is None and shutil.disk_usage(tempfile.gettempdir()).free < 1.5 * bed_bytes:
        return {"unavailable": "no local storage for the synthetic .bed"}
    outdir = tempfile.mkdtemp(prefix="pyrhe_bench_", dir=root)
    try:
        t0 = time.perf_counter()
        paths = write_synthetic_plink(outdir, wl, dev)
        t_write = time.perf_counter() - t0
        import pyrhe.models as models
        from pyrhe.src.util import Logger
        out = {"workload": args.api_workload, "storage": "tmpfs (/dev/shm)" if root else "local disk (page cache after the write)",
               "bed_bytes": bed_bytes, "write_dataset_s": t_write, "runs": []}
        # the first call of the process also pays what a process pays once (kernel module load, the first cudaMalloc of
        # the residency, pinning the staging ring): it is the headline; the same call repeated at the end is `warm_value`
        for cls_name, ring in (("StreamingRHE", None), ("RHE", 4), ("StreamingRHE", 4), ("StreamingRHE", None)):
            if ring is None:
                os.environ.pop("PYRHE_B200_RING_BLOCKS", None)
            else:
                os.environ["PYRHE_B200_RING_BLOCKS"] = str(ring)
            gc.collect()                                       # the previous model's engine (cyclic references) must be gone
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats(dev)
            t0 = time.perf_counter()
            model = getattr(models, cls_name)(model="rhe", num_jack=wl["J"], num_random_vec=wl["B"], seed=0,
                                              geno_impute_method="binary", device="cuda", num_workers=1,
                                              log=Logger(suppress=True, debug_mode=False), **paths)
            t_ctor = time.perf_counter() - t0
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            res = model(trait=0)
            torch.cuda.synchronize(dev)
            t_call = time.perf_counter() - t0
            rep = dict(getattr(model, "ingest_report", {}))
            out["runs"].append({
                "call": f"{cls_name}(...)(trait=0)", "forced_ring_blocks": ring, "constructor_s": t_ctor,
                "call_s": t_call, "value": bed_bytes / t_call / 1e9, "unit": UNIT,
                "staged_gbs": rep.get("staged_bytes", 0) / t_call / 1e9,
                "peak_hbm_gb": torch.cuda.max_memory_allocated(dev) / 1e9, "ingest": rep,
                "sigma_e": float(np.ravel(res["sigma_ests_total"])[-1])})
            del model
        os.environ.pop("PYRHE_B200_RING_BLOCKS", None)
        out["value"] = out["runs"][0]["value"]
        out["warm_value"] = out["runs"][-1]["value"]
        out["unit"] = UNIT
        return out
    finally:
        shutil.rmtree(outdir, ignore_errors=True)


def weighted_shares_fit(J, world, rank, weights, block_bytes, free_bytes, reserve=8e9) -> bool:
    """Whether this rank's share under `weights` (rows + stored partials of its blocks; the individual-major copies adapt
    to what is left) fits the free HBM next to the engine that is already resident."""
    from pyrhe_b200.engine import shard_sizes
    try:
        return bool(shard_sizes(J, world, weights)[rank] * block_bytes + reserve < free_bytes)
    except Exception:
        return False


def build_problem(wl, args, rank, world, dev, shard_weights=None, generate=True):
    """Synthetic inputs of one workload + an engine with this rank's blocks generated in HBM and ingested
    (`generate=False`: the residency is allocated and left for an upload to fill)."""
    import torch
    from pyrhe_b200 import _lib, synth
    from pyrhe_b200.assemble import PathPlan
    from pyrhe_b200.engine import RheEngine, shard_sizes
    from pyrhe_b200.hostmath import host_terms
    lib = _lib.load()
    N, M, J, K, Cc, B = wl["N"], wl["M"], wl["J"], wl["K"], wl["C"], wl["B"]
    rng = np.random.default_rng(0)
    annot = synth.random_annot(M, K, rng)
    np.random.seed(0)
    Z = np.random.randn(N, B)                                 # as base.py:73,176
    W = rng.standard_normal((N, Cc))
    W[:, 0] = rng.random(N) < 0.5
    y = rng.standard_normal((N, 1))
    y -= y.mean()
    plan = PathPlan(model=wl["model"], K=K, B=B, C=Cc, Ty=1)
    env = (rng.random(N) < 0.4).astype(np.float64) if wl["model"] == "genie" else None
    ht, Y_res = host_terms(plan, Z, W, y, env)
    keep = np.ones(N, dtype=bool)

    # memory policy: keep every block's partial in HBM when it fits next to the genotypes
    pitch = (N + 3) // 4
    pitch = (pitch + 127) // 128 * 128
    own_blocks = shard_sizes(J, world, shard_weights)[rank]
    bytes_geno = own_blocks * (M // J + M % J) * pitch
    bytes_part = own_blocks * plan.E * B * pitch * 4 * 4
    free_b, total_b = torch.cuda.mem_get_info(dev)
    store = bytes_geno + bytes_part + 6e9 < free_b
    eng = RheEngine(plan, n_indv=N, keep=keep, annot=annot, num_jack=J, impute="binary", seed=0, device=dev,
                    kernel_path=args.kernel_path, rank=rank, world=world, store_partials=store,
                    retile=getattr(args, "retile", True), shard_weights=shard_weights)
    eng.set_rhs(Z, W, Y_res, env)
    eng.alloc_genotypes()
    if not generate:
        return dict(eng=eng, plan=plan, ht=ht, Z=Z, W=W, Y_res=Y_res, env=env, store=store, ingest_count_ms=None)
    stream = torch.cuda.current_stream(dev)
    for j in eng.own:                                          # synthetic genotypes generated in HBM
        rows, m = eng.block_view(j)
        _lib.check(lib.rhe_synth_genotypes(C.c_void_p(rows.data_ptr()), m, eng.pitch, N, eng.ranges[j][0], 1234, 0.0,
                                           C.c_void_p(stream.cuda_stream)))
    torch.cuda.synchronize(dev)
    # ingest: per-SNP allele counts, once per resident block (what `upload_block` does on arrival)
    ce0, ce1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ce0.record(stream)
    eng.count_all()
    ce1.record(stream)
    torch.cuda.synchronize(dev)
    ingest_count_ms = ce0.elapsed_time(ce1) / max(len(eng.own), 1)          # allele counts + individual-major copies, once per block
    return dict(eng=eng, plan=plan, ht=ht, Z=Z, W=W, Y_res=Y_res, env=env, store=store, ingest_count_ms=ingest_count_ms)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config5", choices=list(WORKLOADS))
    ap.add_argument("--kernel_path", type=int, default=int(os.environ.get("PYRHE_B200_PATH", "1")),
                    help="1 = int8 tcgen05 kernels (default), 0 = CUDA-core validation kernels")
    ap.add_argument("--cpu_snps", type=int, default=200, help="SNPs per block of the CPU sample (oracle port)")
    ap.add_argument("--ref_snps", type=int, default=200, help="SNPs per block of the CPU sample (unmodified reference)")
    ap.add_argument("--cpu_port", action="store_true", help="time the oracle port even when oracle/_ref is present")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_e2e", action="store_true")
    ap.add_argument("--ring_blocks", type=int, default=4, help="distinct host blocks served cyclically in the e2e leg")
    ap.add_argument("--e2e_source", default="pinned", choices=["pinned", "pageable"])
    ap.add_argument("--no_other_configs", action="store_true")
    ap.add_argument("--no_weighted_e2e", action="store_true",
                    help="multi-GPU e2e leg: equal block shares only (default: also shares proportional to the measured H2D rates)")
    ap.add_argument("--force_e2e_weights", default=None, help="comma-separated shard weights for the second e2e leg (testing)")
    ap.add_argument("--no_api_e2e", action="store_true")
    ap.add_argument("--no_retile", dest="retile", action="store_false",
                    help="keep the PLINK rows of blocks that own an individual-major copy (default: re-tile them for pass A)")
    ap.add_argument("--api_workload", default="config2", choices=list(WORKLOADS))
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, wl, rank)
        return

    import torch
    import torch.distributed as dist
    from pyrhe_b200 import _lib
    from pyrhe_b200.assemble import PathPlan, normal_equations_prepare, normal_equations_finish, loo_grams
    from pyrhe_b200.engine import RheEngine, shard_sizes
    from pyrhe_b200.hostmath import host_terms
    from pyrhe_b200 import synth

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1:
        # one process per GPU: host threads and pinned staging memory on the GPU's own NUMA node
        from pyrhe_b200.util.numa import bind_to_gpu_node
        numa = bind_to_gpu_node(local)
        dist.init_process_group("nccl", device_id=dev)
        allnuma = [None] * world
        dist.all_gather_object(allnuma, numa)
        numa = allnuma
    lib = _lib.load()

    N, M, J, K, Cc, B = wl["N"], wl["M"], wl["J"], wl["K"], wl["C"], wl["B"]
    pb = build_problem(wl, args, rank, world, dev)
    eng, plan, ht, Z, W, Y_res, env, store = (pb[k] for k in ("eng", "plan", "ht", "Z", "W", "Y_res", "env", "store"))
    ingest_count_ms = pb["ingest_count_ms"]
    stream = torch.cuda.current_stream(dev)

    def gram_terms(G_blk):
        # host half of the tail that needs only the per-bin Gram pieces: runs while the device still reduces S and
        # forms the leave-one-out Grams (RheEngine.run(gram_hook=...), as Base.pre_compute calls it)
        gram_terms.buf = loo_grams(G_blk, getattr(gram_terms, "buf", None))
        return normal_equations_prepare(plan, ht, gram_terms.buf, eng.Mjk)

    def tail(pieces):
        T, q = normal_equations_finish(pieces["gram_hook"], pieces["XX"])
        return np.linalg.solve(T, q[..., None])[..., 0]

    def step_resident():
        return tail(eng.run(gram_hook=gram_terms))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 1e3 * wall)               # host tail included either way
        timed.local_ms = ms
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    # clocks are sampled (200 ms period) from the first warm-up step to the end of the timed region: at 8 GPUs the
    # timed steps alone are shorter than one sampling period
    with ClockSampler(local) as clk:
        first_pass_ms, sigma = timed(step_resident, 1)         # cold: the first pass of this context over every block
        for _ in range(max(args.warmup - 1, 0)):
            sigma = step_resident()
        launches0 = eng.launches
        ms_total, sigma = timed(step_resident, args.steps)
        launches = (eng.launches - launches0)
        # at eight GPUs warm-up + timed steps last about 0.1 s, less than nvidia-smi needs to deliver its first sample:
        # keep the same step running (untimed) until the sampler has seen the GPU under this load at least three times
        t_extra, clk_extended = time.perf_counter(), False
        while True:
            need = int(len(clk.rows) < 3 and time.perf_counter() - t_extra < 3.0)
            if world > 1:                                      # every rank runs the same number of (collective) steps
                flag = torch.tensor([need], dtype=torch.int32, device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
                need = int(flag.item())
            if not need:
                break
            step_resident()
            clk_extended = True
    ms_step = ms_total / args.steps
    total_bytes = float((N + 3) // 4) * M
    value = total_bytes / (ms_step * 1e-3) / 1e9
    # the round-1 structure for comparison: allele counts re-taken inside every block call (three reads per block)
    ms_recount = None
    if not eng._tiled:                                         # (re-tiled rows cannot be re-counted: --no_retile runs this leg)
        eng.use_resident_counts = False
        step_resident()
        ms_recount, _ = timed(step_resident, 1)
        eng.use_resident_counts = True

    # ---- roofline of the dominant kernel: per-phase CUDA-event timing over one more step
    _lib.check(lib.rhe_timing_enable(eng._ctx, 1))
    step_resident()
    phases = (C.c_double * 4)()
    ncalls = C.c_int32()
    _lib.check(lib.rhe_timing_collect(eng._ctx, phases, C.byref(ncalls)))
    _lib.check(lib.rhe_timing_enable(eng._ctx, 0))
    names = ["params", "pass_a", "standardize_gram", "pass_b"]
    ph = {n: phases[i] / max(ncalls.value, 1) for i, n in enumerate(names)}
    # pass B has two kernels: blocks that own an individual-major copy (as many as the HBM holds next to the SNP-major rows)
    # feed the tensor cores from tensor memory (k_tc_pass_b2); the others gather SNP rows through shared memory (k_tc_pass_b)
    n_own, n_fast = max(len(eng.own), 1), len(eng.gt)
    pb_kernels = {}
    pass_a_kernels = {}
    if args.kernel_path == 1 and len(eng.own) > 0:
        S_buf, P_buf = eng.reserve_state()
        gscr = torch.zeros((plan.E_reg, plan.Rs, plan.Rs), dtype=torch.float64, device=dev)

        def phases_of(blocks):
            _lib.check(lib.rhe_timing_enable(eng._ctx, 1))
            for j in blocks:
                eng._accumulate(j, P_buf[eng.own.index(j)] if P_buf is not None else None, S_buf, gscr)
            out, n = (C.c_double * 4)(), C.c_int32()
            _lib.check(lib.rhe_timing_collect(eng._ctx, out, C.byref(n)))
            _lib.check(lib.rhe_timing_enable(eng._ctx, 0))
            return [out[i] / max(n.value, 1) for i in range(4)]

        fast_blocks = [j for j in eng.own if j in eng.gt]
        slow_blocks = [j for j in eng.own if j not in eng.gt]
        for name_a, name_b, blocks in (
                ("k_tc_pass_a<., 1> (re-tiled rows: contiguous boxes of imputed counts)" if eng._tiled else
                 "k_tc_pass_a<., 0> (PLINK rows)",
                 "k_tc_pass_b2 (A operand from tensor memory, individual-major copy)", fast_blocks),
                ("k_tc_pass_a<., 0> (PLINK rows)", "k_tc_pass_b (SNP rows gathered through shared memory)", slow_blocks)):
            if blocks:
                p4 = phases_of(blocks)
                pass_a_kernels[name_a] = {"blocks": len(blocks), "launch_ms": p4[1]}
                pb_kernels[name_b] = {"blocks": len(blocks), "launch_ms": p4[3]}
        del phases_of, S_buf, P_buf, gscr                      # (views of the engine's accumulators must not outlive it)
    dom = max(("pass_a", "pass_b"), key=lambda n: ph[n])
    m_avg = sum(eng.ranges[j][1] - eng.ranges[j][0] for j in eng.own) / max(len(eng.own), 1)
    alg_bytes = float((N + 3) // 4) * m_avg
    peak, peak_src = load_peaks()
    achieved = alg_bytes / (ph[dom] * 1e-3) / 1e9 if ph[dom] > 0 else 0.0
    for v in list(pb_kernels.values()) + list(pass_a_kernels.values()):
        v["achieved"] = alg_bytes / (v["launch_ms"] * 1e-3) / 1e9
        v["frac"] = v["achieved"] / peak
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath) and args.kernel_path == 1 and args.workload == "config5":
        tr = json.load(open(tpath))
        if dom == "pass_a":                                    # per launch, averaged over this rank's blocks
            n_t = len(eng._tiled)
            traffic = (tr["k_tc_pass_a_retiled"] * n_t + tr["k_tc_pass_a"] * (n_own - n_t)) / n_own
        elif tr.get("k_tc_pass_b2") and tr.get("k_tc_pass_b"):     # per launch, averaged over this rank's blocks
            traffic = (tr["k_tc_pass_b2"] * n_fast + tr["k_tc_pass_b"] * (n_own - n_fast)) / n_own
    block_ms = sum(ph.values())
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": ph[dom], "phases_ms_per_block": ph,
                "pass_a_kernels": pass_a_kernels, "pass_b_kernels": pb_kernels,
                "ingest_ms_per_block": ingest_count_ms,
                "fused_block_frac": alg_bytes / (block_ms * 1e-3) / 1e9 / peak if block_ms > 0 else 0.0}

    # ---- end to end through the engine's ingest API: every block's rows go host memory -> staging threads -> pinned
    # ring -> PCIe -> device slot (+ allele counts on arrival) inside the step.  The host source holds `ring_blocks`
    # distinct blocks that are served cyclically (125 GB of host RAM per rank is not assumed); every byte of the job
    # still crosses the host staging copy and the link each step.
    e2e = None
    h2d_ceiling = None

    def measure_h2d_ceiling():
        """The node's pinned-H2D ceiling with all ranks copying at once (no staging, no kernels): what the links give."""
        pin = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
        dst = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        dst.copy_(pin, non_blocking=True)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(8):
            dst.copy_(pin, non_blocking=True)
        c1.record(stream)
        barrier()
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = 8 * (1 << 30) / (c0.elapsed_time(c1) * 1e-3) / 1e9
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        per_rank = [float(x) for x in t.tolist()]
        # a step that hands every rank the same number of blocks ends when the slowest link is done: with unequal
        # links the ceiling of an equal split is n x the slowest rank's rate, not the sum of the rates
        return {"aggregate_gbs": float(sum(per_rank)), "per_rank_gbs": [round(x, 2) for x in per_rank],
                "equal_split_gbs": world * min(per_rank), "n_gpus": world,
                "how": "8 x 1 GiB pinned->device copies per rank, all ranks concurrently; aggregate = sum of the "
                       "ranks' rates, equal_split = n x the slowest rank's rate"}

    def e2e_leg(eng):
        """Every block's rows go host memory -> (staging threads -> pinned ring ->) PCIe -> device slot (+ ingest
        kernels on arrival) inside the step, through `eng.stream_genotypes` + `eng.run`."""
        if len(eng.own) == 0:
            raise RuntimeError("a rank without blocks")
        R = max(1, min(args.ring_blocks, len(eng.own)))
        host = torch.empty((R * eng.max_m, eng.row_bytes), dtype=torch.uint8, pin_memory=True)
        tmp = torch.zeros((eng.max_m, eng.pitch), dtype=torch.uint8, device=dev)
        for r in range(R):                                     # the same generator as the resident leg (the resident rows may be re-tiled)
            j = eng.own[r]
            m = eng.ranges[j][1] - eng.ranges[j][0]
            _lib.check(lib.rhe_synth_genotypes(C.c_void_p(tmp.data_ptr()), m, eng.pitch, N, eng.ranges[j][0], 1234, 0.0,
                                               C.c_void_p(stream.cuda_stream)))
            host[r * eng.max_m: r * eng.max_m + m].copy_(tmp[:m, : eng.row_bytes])
        torch.cuda.synchronize(dev)
        del tmp
        host_np = host.numpy()

        class CyclicRows:
            """rows [a, b) of the rank's share -> the matching rows of the R-block host sample.  `--e2e_source pinned`
            (default, the contract's "inputs in pinned host memory"): the sample is pinned and copied from where it
            lies; `pageable`: plain host memory, every byte also crosses the staging threads' copy into the ring."""
            shape = (M, eng.row_bytes)

            def _base(self, a):
                j = min(a // (M // J), J - 1)
                return ((j - eng.own[0]) % R) * eng.max_m + (a - eng.ranges[j][0])

            def __getitem__(self, sl):
                base = self._base(sl.start)
                return host_np[base: base + (sl.stop - sl.start)]

            if args.e2e_source == "pinned":
                def pinned_rows(self, a, b):
                    base = self._base(a)
                    return host[base: base + (b - a)]

        h2d_total = M * eng.row_bytes + world * eng.R.numel() * 4      # all ranks: every .bed row once + the RHS per rank
        d2h_holder = {}
        streamer = eng.stream_genotypes(CyclicRows(), ring_blocks=eng.ring_blocks)

        def step_e2e():
            eng.set_rhs(Z, W, Y_res, env)
            pieces = eng.run(upload=streamer, gram_hook=gram_terms)
            d2h_holder["n"] = pieces["XX"].nbytes + pieces["G_blk"].nbytes
            return tail(pieces)

        try:
            step_e2e()
            n_e2e = max(1, min(args.steps, 2))
            ms_e2e, sig = timed(step_e2e, n_e2e)
            ms_e2e /= n_e2e
            # per rank: when the last block of the last timed step had landed on the device, and the H2D rate that is
            streamer.join()
            t = torch.zeros(2 * world, dtype=torch.float64, device=dev)
            t[rank] = 1e3 * (streamer.copy_seconds or 0.0)
            t[world + rank] = (sum(eng.ranges[j][1] - eng.ranges[j][0] for j in eng.own) * eng.row_bytes
                               / max(streamer.copy_seconds or 1.0, 1e-9) / 1e9)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return {"value": total_bytes / (ms_e2e * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(h2d_total), "d2h_bytes_per_step": int(d2h_holder["n"]),
                    "blocks_by_rank": shard_sizes(J, world, eng.shard_weights),
                    "upload_ms_by_rank": [round(float(x), 1) for x in t[:world].tolist()],
                    "upload_gbs_by_rank": [round(float(x), 2) for x in t[world:].tolist()],
                    "host_sample_blocks": R, "host_source": args.e2e_source,
                    "staging_threads": 0 if streamer.pinned_source else streamer.n_workers,
                    "sigma_e": float(sig[-1][-1]),
                    "path": "RheEngine.stream_genotypes (host rows -> staging threads -> pinned ring -> H2D -> counts) + run"}
        finally:
            streamer.close()

    if not args.no_e2e and J >= world:
        h2d_ceiling = measure_h2d_ceiling()
        e2e = e2e_leg(eng)
        rates = h2d_ceiling["per_rank_gbs"]
        if args.force_e2e_weights:
            rates = [float(x) for x in args.force_e2e_weights.split(",")]
        weighted_ok = world > 1 and min(rates) > 0 and min(rates) < 0.9 * max(rates) and not args.no_weighted_e2e \
            and J >= 2 * world
        if weighted_ok:                                       # every rank must have room for its share (agreed by all ranks)
            block_bytes = -(-eng.max_m // 128) * 128 * eng.pitch + plan.E * plan.B * eng.Np * 4
            flag = torch.tensor([int(weighted_shares_fit(J, world, rank, rates, block_bytes,
                                                         torch.cuda.mem_get_info(dev)[0]))], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            weighted_ok = bool(flag.item())
        if weighted_ok:
            # unequal links (on this pool's eight-GPU boxes four GPUs get 23 GB/s and four 35 GB/s when all copy at once):
            # an upload-bound pass ends with the slowest link, so the same leg runs once more on an engine whose
            # contiguous block ranges are proportional to the measured rates (`RheEngine(shard_weights=...)`); the
            # work is the same (every block of the job once; the R-block host sample is laid out per rank, so the two legs
            # see differently ordered synthetic genotypes and report their own sigma_e); the faster leg is the line's e2e
            equal = e2e
            pbw = build_problem(wl, args, rank, world, dev, shard_weights=rates, generate=False)
            try:
                weighted = e2e_leg(pbw["eng"])
            finally:
                pbw["eng"].close()
                del pbw
            weighted["shard_weights"] = rates
            e2e = dict(weighted if weighted["value"] > equal["value"] else equal)
            e2e["equal_split"] = {k: equal[k] for k in ("value", "ms_per_step", "blocks_by_rank", "upload_ms_by_rank", "sigma_e")}
            e2e["rate_weighted_split"] = {k: weighted[k] for k in ("value", "ms_per_step", "blocks_by_rank", "upload_ms_by_rank", "sigma_e")}
        e2e["frac_of_h2d_ceiling"] = e2e["value"] / h2d_ceiling["aggregate_gbs"]
        e2e["frac_of_equal_split_ceiling"] = e2e["value"] / h2d_ceiling["equal_split_gbs"]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        def measure():
            cb, kind = cpu_arm(wl, args)
            try:
                if kind == "port":
                    cb.step()                               # (the port's pool warms up; a reference step is half a minute)
                secs, _, geno = cb.step()
            finally:
                cb.close()
            return cb, kind, secs, geno

        try:
            cb, kind, secs, geno = measure()
        except Exception as exc:
            if args.cpu_port:
                raise
            sys.stderr.write(f"cpu_baseline: the unmodified reference failed ({exc}); timing the oracle port instead\n")
            args.cpu_port = True
            cb, kind, secs, geno = measure()
        cpu = {"value": geno / 4 / secs / 1e9, "unit": UNIT, "cores": cb.cores, "kind": kind,
               "sample": cb.describe(), "sample_seconds": secs}

    eng.close()
    del eng, pb
    torch.cuda.empty_cache()

    # ---- BASELINE.json's other GPU configurations at the current GPU count (resident step, 3 timed steps each)
    others = {}
    if args.workload == "config5" and not args.no_other_configs:
        for name in ("config2", "config3", "config4"):
            owl = dict(WORKLOADS[name])
            opb = build_problem(owl, args, rank, world, dev)
            oeng, oplan, oht = opb["eng"], opb["plan"], opb["ht"]

            def ostep():
                pieces = oeng.run(gram_hook=lambda G: normal_equations_prepare(oplan, oht, loo_grams(G), oeng.Mjk))
                Tq = normal_equations_finish(pieces["gram_hook"], pieces["XX"])
                return np.linalg.solve(Tq[0], Tq[1][..., None])[..., 0]

            ostep()
            oms, osig = timed(ostep, 3)
            oms /= 3
            ob = float((owl["N"] + 3) // 4) * owl["M"]
            others[name] = {"model": owl["model"], "N": owl["N"], "M": owl["M"], "ms_per_step": oms,
                            "value": ob / (oms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
                            "kernel_path": "tcgen05" if oeng.kernel_path == 1 else "simt",
                            "sigma_e": float(osig[-1][-1])}
            oeng.close()
            del oeng, opb
            torch.cuda.empty_cache()

    api = None
    if world == 1 and not args.no_e2e and not args.no_api_e2e:
        api = model_api_e2e(args, dev)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "wall_time_s": ms_step * 1e-3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int8xint8->s32" if args.kernel_path == 1 else "f32",
            "data": "synthetic",
            "config": {"workload": args.workload, **wl, "jackknife_policy": "stored partials" if store else "recompute (streaming)",
                       "kernel_path": "tcgen05" if args.kernel_path == 1 else "simt",
                       "l2": "inputs larger than L2 (packed genotypes per rank >> 126 MB)", "impute": "binary"},
            "e2e": e2e, "gpu_launches": int(launches),
            "clocks": dict(clk.summary(), window="warm-up + timed steps" + (" + the same step repeated untimed until three "
                                                                          "samples were in" if clk_extended else "")),
            "roofline": roofline,
            "cpu_baseline": cpu, "sigma_check": [float(v) for v in sigma[-1]],
            "first_pass_ms": first_pass_ms, "step_recount_ms": ms_recount, "h2d_ceiling": h2d_ceiling,
            "e2e_model_api": api, "other_configs": others, "numa": numa,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _emit(line: dict):
    """The one JSON line goes to the process's original stdout; everything else a library prints while the
    benchmark runs (NCCL's version banner, for one) has been redirected to stderr by then."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # C-level stdout of this process (and of its children) -> stderr
    main()
