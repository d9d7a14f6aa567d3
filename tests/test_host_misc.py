"""CPU: small pieces of host behaviour the drop-in promises -- the lock-file protocol shared with the reference's Locker
(Locker.py:32-69) and the options the reference itself cannot run (mcmc_iters=0 in the MCMC choosers, OPT:316-318, PSEC:254)."""
import os

import numpy as np
import pytest

from tests.helpers import load
from tests.oracle_backend import OracleBackend


def test_lock_files_follow_the_reference_protocol(tmp_path):
    from spearmint_b200.locker import Locker
    f = str(tmp_path / "chooser.pkl")
    a, b = Locker(), Locker()
    assert a.lock(f)
    assert os.path.islink(f + ".lock") and os.readlink(f + ".lock") == "/dev/null"     # what the reference creates / expects
    assert not b.lock(f)                          # another process (or the reference itself) is excluded
    assert a.lock(f)                              # re-entrant for the holder
    assert a.unlock(f) and os.path.islink(f + ".lock")      # still held once
    assert a.unlock(f) and not os.path.lexists(f + ".lock")
    assert b.lock(f) and b.unlock(f)              # free again
    assert a.unlock(f)                            # unlocking what is not held is a no-op, as in the reference
    b.lock_wait(f)                                # returns at once on a free file
    assert os.path.islink(f + ".lock")
    del b                                         # the destructor releases whatever is still held (Locker.py:36-40)
    assert not os.path.lexists(f + ".lock")


@pytest.mark.parametrize("modname", ["GPEIOptChooserB200", "GPEIperSecChooserB200"])
def test_mcmc_iters_zero_is_refused_where_the_reference_is_broken(modname, tmp_path):
    """The reference's mcmc_iters=0 branch of these two choosers dies inside numpy (mismatched arguments / an attribute that
    is never set); there is no behaviour to reproduce, so the plugin says so instead of proposing something."""
    import importlib
    mod = importlib.import_module("spearmint_b200.chooser." + modname)
    g = load("opt_branin2d")
    ch = mod.init(str(tmp_path), "mcmc_iters=0")
    ch._backend = OracleBackend()
    np.random.seed(0)
    with pytest.raises(NotImplementedError):
        ch.next(g["grid"], g["values"], g["durations"], g["candidates"], g["pending"], g["complete"])
    assert ch._backend.loglik_calls == 0          # refused before any work was queued
