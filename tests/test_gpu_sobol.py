"""-m gpu: the device Sobol generator (row f4) against golden points of the REAL reference generator
(tests/golden/sobol.npz, frozen by tools/make_sobol_table.py from sobol_lib.i4_sobol_generate) and, at the grid sizes the
BASELINE configs use, against the oracle restatement.  Integer work: bit-exact (every point is k * 2^-30)."""
import numpy as np
import pytest

from oracle import sobol_oracle as SO
from tests.helpers import load

pytestmark = pytest.mark.gpu


def test_matches_reference_generator():
    from spearmint_b200 import sobol
    g = load("sobol")
    for i in range(7):
        m, n, skip = (int(x) for x in g["case%d_args" % i])
        got = sobol.i4_sobol_generate(m, n, skip)
        assert got.shape == (m, n) and got.dtype == np.float64
        assert np.array_equal(got, g["case%d_pts" % i]), (m, n, skip)


@pytest.mark.parametrize("dims,size,seed", [(8, 10512, 1), (32, 104096, 1), (20, 52048, 7)])
def test_grid_sizes_of_the_baseline_configs(dims, size, seed):
    """N + M points of C2 / headline / C3 as ExperimentGrid would build them; rows checked against the oracle on a
    strided subset, all rows through the Gray-code property: consecutive points differ by exactly one direction number."""
    import torch
    from spearmint_b200 import sobol
    dev = sobol.sobol_device(dims, size, seed)
    grid = dev.cpu().numpy()
    assert grid.shape == (size, dims) and grid.min() >= 0.0 and grid.max() < 1.0
    pick = np.unique(np.concatenate([np.arange(0, size, 997), [size - 1]]))
    V = SO.direction_numbers()[:dims].astype(np.int64)
    for j in pick:
        s = max(seed + int(j) - 1, 0)
        gcode = s ^ (s >> 1)
        q = np.zeros(dims, dtype=np.int64)
        b = 0
        while gcode:
            if gcode & 1:
                q ^= V[:, b]
            gcode >>= 1
            b += 1
        assert np.array_equal(grid[j], q * 2.0 ** -30), j
    ints = np.rint(grid * 2.0 ** 30).astype(np.int64)
    assert np.array_equal(ints * 2.0 ** -30, grid)
    s0 = max(seed - 1, 0)
    for j in range(1, min(size, 5000)):                      # lastq(s+1) = lastq(s) ^ v[:, lo0(s)]
        s = s0 + j - 1
        lo0 = (~s & (s + 1)).bit_length() - 1
        assert np.array_equal(ints[j] ^ ints[j - 1], V[:, lo0]), j
    f32 = sobol.sobol_device(dims, 4096, seed, dtype=torch.float32).double().cpu().numpy()
    assert np.abs(f32 - grid[:4096]).max() <= 2.0 ** -24


def test_limits():
    from spearmint_b200 import sobol
    with pytest.raises(ValueError):
        sobol.sobol_device(1112, 4, 1)
    with pytest.raises(Exception):
        sobol.sobol_device(2, 8, 2 ** 30)                     # the reference prints "Too many calls!" there
