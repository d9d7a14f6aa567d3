"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's Sobol generator (parity oracle of row f4).

Follows ``spearmint/spearmint/sobol_lib.py``: ``i4_sobol_generate`` (:125-156) calls ``i4_sobol`` (:158-13787) with
seeds skip-1, skip, ...; ``i4_sobol`` is the Antonov-Saleev Gray-code recurrence
    lastq(seed + 1) = lastq(seed) XOR v[:, lo0(seed)],   point(seed) = lastq(seed) * 2^-30,   lastq(0) = 0,
(lo0 = position of the lowest zero bit, ``i4_bit_lo0`` :72-122), negative seeds clamped to 0.  Unrolled:
    lastq(seed) = XOR over the set bits b of gray(seed) = seed ^ (seed >> 1) of v[:, b].
The direction numbers v are the ones the reference builds from its Joe-Kuo table (1111 x 30 integers, frozen from the
executing reference by tools/make_sobol_table.py).  Pinned against golden points of the real generator
(tests/golden/sobol.npz, tests/test_oracle_golden.py).  Only tests/ may import this module.
"""
import os

import numpy as np

_V = None


def direction_numbers():
    global _V
    if _V is None:
        here = os.path.dirname(os.path.abspath(__file__))
        _V = np.load(os.path.join(here, "..", "spearmint_b200", "data", "sobol_v_1111x30.npy"))
    return _V


def i4_sobol_generate(m, n, skip):
    """(m, n) float64 array, column j = point of seed max(skip + j - 1, 0) (sobol_lib.py:152-155)."""
    V = direction_numbers()[:m].astype(np.int64)
    out = np.zeros((m, n))
    for j in range(n):
        seed = max(skip + j - 1, 0)
        g = seed ^ (seed >> 1)
        q = np.zeros(m, dtype=np.int64)
        b = 0
        while g:
            if g & 1:
                q ^= V[:, b]
            g >>= 1
            b += 1
        out[:, j] = q * 2.0 ** -30
    return out
