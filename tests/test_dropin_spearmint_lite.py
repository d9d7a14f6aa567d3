"""CPU, build container only (needs /root/reference): BASELINE config 1 -- the reference's UNMODIFIED spearmint-lite
controller proposes Branin experiments through our chooser, loaded by name via the 3-line stub of INTEGRATION.md.
The controller, GridMap, results.dat parsing and candidate generation are the reference's own code (py3 shim);
only the numerics behind the chooser are the oracle stand-in here (no GPU in this container) -- the same chooser
class is exercised with the real DeviceBackend in tests/test_gpu_chooser.py."""
import os
import shutil
import sys
import types

import numpy as np
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")

STUB = "from spearmint_b200.chooser.GPEIOptChooserB200 import *  # noqa: F401,F403\n"


def branin(x):
    """spearmint/examples/braninpy/branin.py:6-16 (x in the unit square)."""
    x1, x2 = x[0] * 15 - 5, x[1] * 15
    return float(np.square(x2 - (5.1 / (4 * np.square(np.pi))) * np.square(x1) + (5 / np.pi) * x1 - 6)
                 + 10 * (1 - (1. / (8 * np.pi))) * np.cos(x1) + 10)


def test_spearmint_lite_drives_the_dropin_chooser(tmp_path, monkeypatch):
    lite = ref_shim.load_lite()
    # the stub a maintainer drops into spearmint/spearmint/chooser/
    plug = tmp_path / "plugins"
    plug.mkdir()
    (plug / "GPEIOptChooserB200.py").write_text(STUB)
    sys.modules["chooser"].__path__ = [str(plug)]
    sys.modules.pop("chooser.GPEIOptChooserB200", None)
    # numerics: oracle stand-in (CPU container); the product default is DeviceBackend and raises without CUDA
    import spearmint_b200.backend as be
    from tests.oracle_backend import OracleBackend
    monkeypatch.setattr(be, "DeviceBackend", lambda **kw: OracleBackend())

    expt = tmp_path / "braninpy"
    expt.mkdir()
    shutil.copy(os.path.join(ref_shim.REF_ROOT, "spearmint-lite", "braninpy", "config.json"), str(expt))
    opts = types.SimpleNamespace(num_jobs=1, max_finished_jobs=1000, chooser_module="GPEIOptChooserB200",
                                 chooser_args="mcmc_iters=3,burnin=5,grid_subset=4,use_multiprocessing=0",
                                 grid_size=300, grid_seed=1, config_file="config.json", results_file="results.dat")
    res = expt / "results.dat"
    np.random.seed(0)
    best = np.inf
    for it in range(8):
        lite.main_controller(opts, [str(expt)])            # appends one "P P x1 x2" line (LITE:199-215)
        lines = res.read_text().splitlines()
        assert lines[-1].startswith("P P ")
        x = [float(t) for t in lines[-1].split()[2:]]
        assert len(x) == 2 and all(0.0 <= t <= 1.0 for t in x)
        y = branin(x)
        best = min(best, y)
        lines[-1] = "%f 1 %s" % (y, " ".join(lines[-1].split()[2:]))      # braninrunner.py fills the result in
        res.write_text("\n".join(lines) + "\n")
    assert len(res.read_text().splitlines()) == 8
    # GP-EI picked the last 6 (the first two proposals are "candidates[0]" with < 2 completed, OPT:221-222)
    assert os.path.exists(str(expt / "spearmint_b200.chooser.GPEIOptChooserB200.pkl"))
    assert best < 25.0      # Branin over the unit square spans 0.4 .. 300; 8 evaluations already land in the basin
