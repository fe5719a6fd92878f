"""Generates tests/golden/sod_projection.npz with the REAL reference projection code
(oracle/_ref, built from /root/reference by oracle/Makefile): inputs and the outputs of
project_to_SO3 / project_to_SO2 (C++/DPGO/src/internal/project_to_SOd.cpp) compiled with
-ffp-contract=off, i.e. exactly the mul / add / fma mix the source spells.
Run in the authoring container:  python tests/golden/make_projection_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref  # noqa: E402
import test_oracle_ref as T  # noqa: E402

out = {}
for d in (2, 3):
    A = T._inputs(d, np.random.default_rng(100 + d))[:450]      # the non-degenerate blocks
    out["A%d" % d] = A
    out["U%d" % d] = ref.project(A, nofma=True)
    out["U%d_fma" % d] = ref.project(A)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref", "sod_projection.npz"), **out)
print("wrote", {k: v.shape for k, v in out.items()})
