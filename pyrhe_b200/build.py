"""Build libpyrhe_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libpyrhe_b200.so")
SOURCES = ["rhe_capi.cu", "rhe_tc.cu"]
HEADERS = ["rhe_common.cuh", os.path.join("..", "..", "include", "pyrhe_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, out=None):
    """`out`: write another build of the library there (ablation / profiling builds with PYRHE_B200_EXTRA_NVCC)."""
    global LIB
    if out:
        LIB, force = out, True
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("PYRHE_B200_EXTRA_NVCC", "").split()      # e.g. -DRHE_TC_DEBUG -DRHE_TC_PROF (profiling builds)
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES] + ["-lcuda"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libpyrhe_b200.so")
    with open(os.path.join(CSRC, "ptxas_report.txt"), "w") as f:
        f.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    out = [a for a in sys.argv[1:] if a.endswith(".so")]
    print(build(force="--force" in sys.argv, verbose=True, out=out[0] if out else None))
