"""fp64 yardstick for the N-scale parity cases:  python tools/make_scale_yardstick.py [case ...]

Runs the CPU oracle with float64 block products (oracle.rhe_oracle.run(f64=True): the reference's algorithm without
its fp32 rounding) on the inputs of every scale case and stores T, q, sigma^2 and SEs in tests/golden/<case>.fp64.npz.
tests/test_gpu_scale.py uses them to state, next to the error of the CUDA path against the reference, how far the
reference's own fp32 arithmetic is from the exact answer on the same inputs -- the envelope a parity tolerance can
meaningfully ask for (SURVEY.md §9.2 measured it at N = 5000 only).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from golden_cases import SCALE_CASES  # noqa: E402


def main():
    import helpers
    from oracle import rhe_oracle
    for name in sys.argv[1:] or SCALE_CASES:
        p = helpers.oracle_problem(name)
        out = rhe_oracle.run(p, f64=True)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".fp64.npz"), T=out["T"], q=out["q"],
                            sigma_total=out["sigma_total"], sigma_jack=out["sigma_jack"], sigma_se=out["sigma_se"])
        print(name, "fp64 sigma:", out["sigma_total"])


if __name__ == "__main__":
    main()
