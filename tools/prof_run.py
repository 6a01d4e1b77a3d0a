"""Per-role cycle accounting of the tensor kernels (where decode / MMA-issue warps wait): builds a `-DRHE_TC_PROF` copy of
the library next to the shipped one if it is missing, runs four config-5 blocks through it and prints the counters.

    python tools/prof_run.py            (on the GPU box; the build needs nvcc)
"""
import os
import runpy
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PROF = os.path.join(ROOT, "pyrhe_b200", "csrc", "libpyrhe_b200_prof.so")
if not os.path.exists(PROF):
    env = dict(os.environ, PYRHE_B200_EXTRA_NVCC="-DRHE_TC_PROF")
    subprocess.run([sys.executable, os.path.join(ROOT, "pyrhe_b200", "build.py"), PROF], env=env, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([sys.executable, os.path.join(ROOT, "pyrhe_b200", "build.py"), "--force"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)      # (restores the ptxas report of the shipped build)
os.environ["PYRHE_B200_LIB"] = PROF
from pyrhe_b200 import _lib  # noqa: E402

sys.argv = ["bench.py", "--workload", "profile5", "--steps", "1", "--warmup", "0", "--no_cpu_baseline", "--no_e2e",
            "--no_api_e2e", "--no_retile"]
try:
    runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
finally:
    _lib.load().rhe_tc_prof_dump()
