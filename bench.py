"""Benchmark of the RHE trace-estimation hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 3 --warmup 3                       # this implementation
    python -m torch.distributed.run --nproc-per-node 8 ... bench.py --gpus 8 ...
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1      # CPU arm (oracle port)

A "step" is one full RHE jackknife over synthetic genotypes of the named shape: every jackknife
block of every rank through the block kernels, the all-reduce of the totals, the leave-one-out
Grams, the D2H of the Gram pieces, host assembly of the J+1 normal equations and their solves.
`value` = packed .bed bytes (ceil(N0/4) * M) / step time with the genotypes resident in HBM;
`e2e` = the same step with every block's rows copied from pinned host memory inside the timed
region.  The problem size is fixed as GPUs are added ("scaling": "strong"), as BASELINE.json
quotes the metric for one problem at 1/2/4/8 GPUs.
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
def reference_arm(args, wl, rank):
    """CPU implementation of the path (oracle port), all host cores, bounded sample per step."""
    if rank != 0:
        return
    from oracle.cpu_baseline import CpuBaseline
    cb = CpuBaseline(wl["N"], wl["K"], wl["C"], wl["B"], snps_per_block=args.cpu_snps)
    for _ in range(args.warmup):
        cb.step()
    secs, geno = 0.0, 0.0
    for _ in range(args.steps):
        s, _, g = cb.step()
        secs += s
        geno += g
    cb.close()
    gbs = geno / 4 / secs / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, **wl, "sample": cb.describe()},
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": cb.cores, "kind": "port", "sample": cb.describe()},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "projected_full_job_s": wl["N"] / 4 * wl["M"] / 1e9 / gbs,
    }
    _emit(line)


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config5", choices=list(WORKLOADS))
    ap.add_argument("--kernel_path", type=int, default=int(os.environ.get("PYRHE_B200_PATH", "1")),
                    help="1 = int8 tcgen05 kernels (default), 0 = CUDA-core validation kernels")
    ap.add_argument("--cpu_snps", type=int, default=200, help="SNPs per block of the CPU sample")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_e2e", action="store_true")
    ap.add_argument("--ring_blocks", type=int, default=4, help="pinned host ring (blocks) for the e2e leg")
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
    from pyrhe_b200.assemble import PathPlan, normal_equations_batch, loo_grams
    from pyrhe_b200.engine import RheEngine
    from pyrhe_b200.hostmath import host_terms
    from pyrhe_b200 import synth

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
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
    per = -(-J // world)
    own_blocks = max(0, min((rank + 1) * per, J) - rank * per)
    bytes_geno = own_blocks * (M // J + M % J) * pitch
    bytes_part = own_blocks * plan.E * B * pitch * 4 * 4
    free_b, total_b = torch.cuda.mem_get_info(dev)
    store = bytes_geno + bytes_part + 6e9 < free_b
    eng = RheEngine(plan, n_indv=N, keep=keep, annot=annot, num_jack=J, impute="binary", seed=0, device=dev,
                    kernel_path=args.kernel_path, rank=rank, world=world, store_partials=store)
    eng.set_rhs(Z, W, Y_res, env)
    eng.alloc_genotypes()
    stream = torch.cuda.current_stream(dev)
    for j in eng.own:                                          # synthetic genotypes generated in HBM
        rows, m = eng.block_view(j)
        _lib.check(lib.rhe_synth_genotypes(C.c_void_p(rows.data_ptr()), m, eng.pitch, N, eng.ranges[j][0], 1234, 0.0,
                                           C.c_void_p(stream.cuda_stream)))
    torch.cuda.synchronize(dev)

    def tail(pieces):
        tail.buf = loo_grams(pieces["G_blk"], getattr(tail, "buf", None))
        T, q = normal_equations_batch(plan, ht, pieces["XX"], tail.buf, pieces["M"])
        try:
            return np.linalg.solve(T, q[..., None])[..., 0]
        except np.linalg.LinAlgError:        # only with the kernel debug switches (PYRHE_TC_DEBUG_*) that skip work
            return np.full(q.shape, np.nan)

    def step_resident():
        return tail(eng.run())

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
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    # clocks are sampled (200 ms period) from the first warm-up step to the end of the timed region: at 8 GPUs the
    # timed steps alone are shorter than one sampling period
    with ClockSampler(local) as clk:
        for _ in range(args.warmup):
            sigma = step_resident()
        launches0 = eng.launches
        ms_total, sigma = timed(step_resident, args.steps)
    launches = (eng.launches - launches0)
    ms_step = ms_total / args.steps
    total_bytes = float((N + 3) // 4) * M
    value = total_bytes / (ms_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel: per-phase CUDA-event timing over one more step
    _lib.check(lib.rhe_timing_enable(eng._ctx, 1))
    step_resident()
    phases = (C.c_double * 4)()
    ncalls = C.c_int32()
    _lib.check(lib.rhe_timing_collect(eng._ctx, phases, C.byref(ncalls)))
    _lib.check(lib.rhe_timing_enable(eng._ctx, 0))
    names = ["stats_impute", "pass_a", "standardize_gram", "pass_b"]
    ph = {n: phases[i] / max(ncalls.value, 1) for i, n in enumerate(names)}
    dom = max(("pass_a", "pass_b"), key=lambda n: ph[n])
    m_avg = sum(eng.ranges[j][1] - eng.ranges[j][0] for j in eng.own) / max(len(eng.own), 1)
    alg_bytes = float((N + 3) // 4) * m_avg
    peak, peak_src = load_peaks()
    achieved = alg_bytes / (ph[dom] * 1e-3) / 1e9 if ph[dom] > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(f"{dom}:{args.kernel_path}:{args.workload}")
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": ph[dom], "phases_ms_per_block": ph,
                "fused_block_frac": alg_bytes / (sum(ph.values()) * 1e-3) / 1e9 / peak if sum(ph.values()) > 0 else 0.0}

    # ---- end to end: every block's rows cross PCIe from pinned host memory inside the step
    e2e = None
    if not args.no_e2e:
        R = max(1, min(args.ring_blocks, len(eng.own)))
        ring = torch.empty((R, eng.max_m, eng.row_bytes), dtype=torch.uint8).pin_memory()
        for r in range(R):
            rows, m = eng.block_view(eng.own[r])
            ring[r, :m].copy_(rows[:, : eng.row_bytes])
        copy_stream = torch.cuda.Stream(dev)
        h2d_total = M * eng.row_bytes + world * eng.R.numel() * 4      # all ranks: every .bed row once + the RHS per rank
        d2h_holder = {}

        def step_e2e():
            eng.set_rhs(Z, W, Y_res, env)
            events = {}
            for idx, j in enumerate(eng.own):
                eng.upload_block(j, ring[idx % R], stream=copy_stream)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                events[j] = ev
            pieces = eng.run(upload_events=events)
            d2h_holder["n"] = pieces["XX"].nbytes + pieces["G_blk"].nbytes
            return tail(pieces)

        step_e2e()
        ms_e2e, _ = timed(step_e2e, max(1, min(args.steps, 2)))
        ms_e2e /= max(1, min(args.steps, 2))
        e2e = {"value": total_bytes / (ms_e2e * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(h2d_total), "d2h_bytes_per_step": int(d2h_holder["n"]),
               "host_ring_blocks": R}
        del ring

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_baseline import CpuBaseline
        cb = CpuBaseline(N, K, Cc, B, snps_per_block=args.cpu_snps)
        cb.step()
        secs, _, geno = cb.step()
        cb.close()
        cpu = {"value": geno / 4 / secs / 1e9, "unit": UNIT, "cores": cb.cores, "kind": "port",
               "sample": cb.describe(), "sample_seconds": secs}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "wall_time_s": ms_step * 1e-3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int8xint8->s32" if args.kernel_path == 1 else "f32",
            "data": "synthetic",
            "config": {"workload": args.workload, **wl, "jackknife_policy": "stored partials" if store else "recompute (streaming)",
                       "kernel_path": "tcgen05" if args.kernel_path == 1 else "simt",
                       "l2": "inputs larger than L2 (packed genotypes per rank >> 126 MB)", "impute": "binary"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roofline,
            "cpu_baseline": cpu, "sigma_check": [float(v) for v in sigma[-1]],
        }
        _emit(line)
    eng.close()
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
