#!/usr/bin/env python
"""Writes spearmint_b200/data/sobol_v_1111x30.npy: the direction-number matrix V (1111 dimensions x 30 bits, uint32)
of the reference's Sobol generator, and tests/golden/sobol.npz: golden points of the reference generator.

The reference's ``sobol_lib.py`` (spearmint/spearmint/sobol_lib.py:158-13787; Burkardt's I4_SOBOL with the Joe-Kuo
primitive polynomials and initial direction numbers for 1111 dimensions) builds V once -- the recurrence over the
primitive polynomial of each dimension, then the scaling v[:, j] *= 2^(29 - j) -- and keeps it in a module global.
This script EXECUTES that code through oracle/ref_shim.py and saves the resulting integers; nothing is copied from
the reference's source.  (V differs from scipy.stats.qmc.Sobol's table for dimensions >= 3, so it cannot be taken from
scipy -- SURVEY.md section 2, row 10.)  Run in the build container (needs /root/reference):  python tools/make_sobol_table.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

sb = ref_shim.load()["sobol_lib"]
sb.i4_sobol(1111, 0)                       # initialises the module globals for the maximum dimension
v = np.array(sb.v)
assert v.shape == (1111, 30) and sb.maxcol == 30 and sb.recipd == 2.0 ** -30
V = v.astype(np.uint32)
assert np.all(V == v) and V.max() < 2 ** 30
out = os.path.join(ROOT, "spearmint_b200", "data", "sobol_v_1111x30.npy")
np.save(out, V)
print("wrote", out, V.shape, V.dtype)

# golden points straight from the reference's generator: (m, n, skip) -> (m, n) array (GRID:192-196 calls it with
# skip = grid_seed and transposes)
cases = [(2, 40, 1), (5, 64, 1), (8, 100, 3), (32, 50, 1000), (40, 33, 1), (1111, 4, 7), (3, 16, 0)]
gold = {}
for i, (m, n, skip) in enumerate(cases):
    gold["case%d_args" % i] = np.array([m, n, skip])
    gold["case%d_pts" % i] = sb.i4_sobol_generate(m, n, skip)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sobol.npz"), **gold)
print("wrote tests/golden/sobol.npz with", len(cases), "cases")
