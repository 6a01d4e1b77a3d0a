"""The `e2e_model_api` leg of bench.py on its own (StreamingRHE / RHE on a real .bed file at config-2 size)."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

torch.cuda.set_device(0)
args = types.SimpleNamespace(api_workload=sys.argv[1] if len(sys.argv) > 1 else "config2", kernel_path=1)
out = bench.model_api_e2e(args, torch.device("cuda", 0))
for r in out["runs"]:
    print(r["call"], "ring", r["forced_ring_blocks"], f"ctor {r['constructor_s']:.2f} s, call {r['call_s']:.3f} s, {r['value']:.1f} GB/s, "
          f"staged {r['staged_gbs']:.1f} GB/s, peak HBM {r['peak_hbm_gb']:.1f} GB", r["ingest"])
