"""Regenerates tests/golden/*.npz with the CPU oracle (run once per variant: plain = `wide` mul_add unfused, the
default; RAYN_MULADD_FUSED=1 writes the *_fma.npz set).  These are SELF-generated pins (the
reference has no golden vectors): they guard the oracle against regressions and give
the GPU tests committed numbers to compare with.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import binding as ob  # noqa: E402
from rayn_b200 import configs  # noqa: E402
from helpers import small_config  # noqa: E402
from test_cpu_oracle import GOLD_SUFFIX, GOLDEN_CASES  # noqa: E402

if __name__ == "__main__":
    for name, (n, res, samples, mb) in GOLDEN_CASES.items():
        c, inp = small_config(n, res, samples, mb)
        o, info = ob.render(c["world"], c["camera"], inp, (16, 16), c["integrator"], configs.frame_time_range(1))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + GOLD_SUFFIX + ".npz"), **o)
        print(name, info)
