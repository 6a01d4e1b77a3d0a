"""Where one step goes when a rank owns 13 config-5 blocks (the 8-GPU share): block kernels, leave-one-out Grams,
device -> host copies, host assembly and solves."""
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from pyrhe_b200.assemble import normal_equations_batch, normal_equations_prepare, normal_equations_finish, loo_grams

wl = dict(N=500_000, M=130_000, J=13, K=8, C=5, B=10, model=sys.argv[1] if len(sys.argv) > 1 else "rhe")
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
pb = bench.build_problem(wl, types.SimpleNamespace(kernel_path=1), 0, 1, dev)
eng, plan, ht = pb["eng"], pb["plan"], pb["ht"]


def tail(pieces, buf=[None]):
    buf[0] = loo_grams(pieces["G_blk"], buf[0])
    T, q = normal_equations_batch(plan, ht, pieces["XX"], buf[0], pieces["M"])
    return np.linalg.solve(T, q[..., None])[..., 0]


def gram_terms(G_blk, buf=[None]):
    buf[0] = loo_grams(G_blk, buf[0])
    return normal_equations_prepare(plan, ht, buf[0], eng.Mjk)


def step_overlapped():
    pieces = eng.run(gram_hook=gram_terms)
    T, q = normal_equations_finish(pieces["gram_hook"], pieces["XX"])
    return np.linalg.solve(T, q[..., None])[..., 0]


n = 10
for name, fn in (("host tail after run()", lambda: tail(eng.run())),
                 ("Gram half of the tail inside run() (gram_hook)", step_overlapped)):
    for _ in range(3):
        ref = fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"{name}: {1e3 * (t1 - t0) / n:.3f} ms per step, sigma_e {out[-1][-1]:.12f}; last run's tail (wait for the Gram "
          f"pieces, gram_hook, wait for XX) ms: {[round(1e3 * x, 3) for x in eng.tail_seconds]}")
    t0 = time.perf_counter()
    for _ in range(n):
        pieces = eng.run()
    t1 = time.perf_counter()
    for _ in range(n):
        tail(pieces)
    t2 = time.perf_counter()
    for _ in range(n):
        g = gram_terms(pieces["G_blk"])
    t3 = time.perf_counter()
    print(f"   run() alone {1e3 * (t1 - t0) / n:.3f} ms, whole host tail {1e3 * (t2 - t1) / n:.3f} ms, its Gram half "
          f"{1e3 * (t3 - t2) / n:.3f} ms")
# inside run(): device time of the block loop alone
S, P_all = eng.reserve_state()
G_blk = torch.zeros((eng.J, plan.E_reg, plan.Rs, plan.Rs), dtype=torch.float64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    eng._pass(None, lambda jl, j: eng._accumulate(j, P_all[jl], S, G_blk[j]))
e1.record()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"block loop: device {e0.elapsed_time(e1) / n:.3f} ms, host enqueue {1e3 * t_host / n:.3f} ms per step")
