#!/usr/bin/env python
"""Generates tests/golden/mll_*.npz by EXECUTING the reference's GP.optimize_hypers (gp.py:181-292) and
GPEIChooser.next() with mcmc_iters=0 (chooser/GPEIChooser.py:124-176, 348-361) through oracle/ref_shim.py.

The objective / gradient closures of optimize_hypers are local to the function, so they are captured by wrapping the
gp module's ``spo.fmin_l_bfgs_b`` for the duration of the call: the wrapper records (f, g) at the start point and at
a few perturbed points, then lets the real optimiser run.  Run in the build container only: python tests/golden/make_golden_mll.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_shim  # noqa: E402

R = ref_shim.load()
gp, GPEI, sobol = R["gp"], R["GPEI"], R["sobol_lib"]


def problem(D, N, M, seed):
    G = N + M
    grid = sobol.i4_sobol_generate(D, G, 1).T
    perm = np.random.RandomState(seed).permutation(G)
    complete, candidates = np.sort(perm[:N]), np.sort(perm[N:])
    values = np.full(G, np.nan)
    y = np.sin(3 * grid[complete]).sum(1) + 0.05 * np.random.RandomState(seed + 1).randn(N)
    values[complete] = (y - y.mean()) / y.std()
    return grid, values, candidates, complete


def one(name, kind, D, N, M, seed):
    grid, values, candidates, complete = problem(D, N, M, seed)
    comp, vals = grid[complete], values[complete]
    rec = {}
    real = gp.spo.fmin_l_bfgs_b

    def wrapped(f, x0, fprime, **kw):
        rs = np.random.RandomState(seed + 7)
        pts = [np.array(x0, dtype=float)] + [np.array(x0) + 0.3 * rs.randn(len(x0)) for _ in range(4)]
        rec["pts"] = np.array(pts)
        rec["f"] = np.array([f(p) for p in pts])
        rec["g"] = np.array([fprime(p) for p in pts])
        out = real(f, x0, fprime, **kw)
        rec["evals"] = out[2]["funcalls"]
        return out

    gp.spo.fmin_l_bfgs_b = wrapped
    try:
        g = gp.GP(kind)
        g.real_init(D, vals)
        g.optimize_hypers(comp, vals)
    finally:
        gp.spo.fmin_l_bfgs_b = real
    out = dict(kind=kind, grid=grid, values=values, candidates=candidates, complete=complete,
               pts=rec["pts"], f=rec["f"], g=rec["g"], evals=rec["evals"],
               opt_mean=g.mean, opt_noise=g.noise, opt_amp2=g.amp2, opt_ls=g.ls)
    # the chooser entry point with mcmc_iters=0
    ch = GPEI.init(tempfile.mkdtemp(), "covar=%s,mcmc_iters=0" % kind)
    np.random.seed(seed)
    ret = ch.next(grid, values, np.full(grid.shape[0], np.nan), candidates, np.array([], dtype=int), complete)
    out["next_index"] = int(ret)
    out["next_hypers"] = np.hstack([ch.mean, ch.noise, ch.amp2, ch.ls])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "evals", rec["evals"], "amp2 %.4f noise %.5f ls" % (g.amp2, g.noise), np.round(g.ls, 3), "next", ret)


one("mll_d3_m52", "Matern52", 3, 40, 300, 5)
one("mll_d5_ardse", "ARDSE", 5, 60, 400, 6)
one("mll_d2_m32", "Matern32", 2, 25, 200, 7)
