"""Drop-in replacement for the reference's ``chooser/GPEIChooser.py`` (EI on the grid only, no refinement).

``next()`` (GPEIChooser.py:124-176): for each of ``mcmc_iters`` iterations draw one hyper-sample (no burn-in, the
chain continues from the previous call / pickle) and evaluate EI over all candidates; return the argmax of the mean.
The reference interleaves RNG use -- sample, fantasy normals (``npr.randn(P, F)``, GPEI:237, no state reset), sample,
... -- so the normals are drawn here in that same order and carried per sample; the EI passes themselves are then
done for all samples in ONE batched GPU pass (``compute_ei`` math is identical to GPEIOptChooser's, GPEI:178-266).
State pickle keys dims/ls/amp2/noise/mean, written on destruction like the reference (GPEI:66-84).
``mcmc_iters=0`` runs the ML-II branch: ``gp.GP.optimize_hypers`` (spearmint_b200/gp.py, GP:181-292) then EI under the optimum.
"""
import os
import pickle
import tempfile

import numpy as np
import numpy.random as npr

from spearmint_b200 import util
from spearmint_b200.locker import Locker, log

COVARS = ("SE", "ARDSE", "Matern32", "Matern52")


def init(expt_dir, arg_string):
    args = util.unpack_args(arg_string)
    return GPEIChooserB200(expt_dir, **args)


class GPEIChooserB200(object):

    def __init__(self, expt_dir, covar="Matern52", mcmc_iters=10, pending_samples=100, noiseless=False,
                 device=None, state_name=None, backend=None):
        if covar not in COVARS:
            raise AttributeError("module 'spearmint.gp' has no attribute '%s'" % covar)
        self.covar = covar
        self.locker = Locker()
        name = state_name if state_name else self.__module__
        self.state_pkl = os.path.join(expt_dir, name + ".pkl")
        self.mcmc_iters = int(mcmc_iters)
        self.pending_samples = int(pending_samples)
        self.D = -1
        self.hyper_iters = 1
        self.noiseless = bool(int(noiseless))
        self.noise_scale, self.amp2_scale, self.max_ls = 0.1, 1, 2
        self._device, self._backend = device, backend
        self._ll = None

    @property
    def backend(self):
        if self._backend is None:
            from spearmint_b200.backend import DeviceBackend
            self._backend = DeviceBackend(device=self._device)
        return self._backend

    def dump_hypers(self):
        if self.D == -1:
            return
        self.locker.lock_wait(self.state_pkl)
        fh = tempfile.NamedTemporaryFile(mode="wb", delete=False)
        pickle.dump({"dims": self.D, "ls": self.ls, "amp2": self.amp2, "noise": self.noise, "mean": self.mean},
                    fh, protocol=2)
        fh.close()
        os.system('mv "%s" "%s"' % (fh.name, self.state_pkl))
        self.locker.unlock(self.state_pkl)

    def __del__(self):          # the reference persists its state in the destructor (GPEI:66-84)
        try:
            self.dump_hypers()
        except Exception:
            pass

    def _real_init(self, dims, values):
        self.locker.lock_wait(self.state_pkl)
        if os.path.exists(self.state_pkl):
            with open(self.state_pkl, "rb") as fh:
                state = pickle.load(fh)
            self.D, self.ls, self.amp2 = state["dims"], state["ls"], state["amp2"]
            self.noise, self.mean = state["noise"], state["mean"]
        else:
            self.D = dims
            self.ls = np.ones(self.D)
            self.amp2 = np.std(values) + 1e-4
            self.noise = 1e-3
            self.mean = np.mean(values)
        self.locker.unlock(self.state_pkl)

    def next(self, grid, values, durations, candidates, pending, complete):
        if complete.shape[0] < 2:
            return int(candidates[0])
        if self.D == -1:
            self._real_init(grid.shape[1], values[complete])
        comp, cand, pend = grid[complete, :], grid[candidates, :], grid[pending, :]
        vals = values[complete]
        if self.mcmc_iters <= 0:
            # ML-II branch (GPEI:156-176): optimise the hyper-parameters, EI under them, argmax
            try:
                self.optimize_hypers(comp, vals)
            except Exception:                         # the reference's bare except: fall back to the initial values
                self.ls = np.ones(self.D)
                self.amp2 = np.std(vals)
                self.noise = 1e-3
            log("mean: %f  amp: %f  noise: %f  min_ls: %f  max_ls: %f"
                % (self.mean, np.sqrt(self.amp2), self.noise, np.min(self.ls), np.max(self.ls)))
            ei = self.compute_ei(comp, pend, cand, vals)
            return int(candidates[int(np.argmax(ei))])
        P = pend.shape[0]
        self._ll = self.backend.loglik(self.covar, comp, vals)
        hs, normals = [], []
        for mcmc_iter in range(self.mcmc_iters):
            self.sample_hypers(comp, vals)
            log("mean: %f  amp: %f  noise: %f  min_ls: %f  max_ls: %f"
                % (self.mean, np.sqrt(self.amp2), self.noise, np.min(self.ls), np.max(self.ls)))
            hs.append((self.mean, self.noise, self.amp2, self.ls))
            if P:
                normals.append(npr.randn(P, self.pending_samples))       # same stream position as GPEI:237
        self._ll = None
        st = self.backend.grid_state(self.covar, hs, comp, pend, vals, np.array(normals) if P else None)
        best_cand = int(self.backend.top_mean_ei(st, cand, 1)[-1])
        return int(candidates[best_cand])

    def compute_ei(self, comp, pend, cand, vals):
        """EI under the current hyper-parameters (GPEI:178-266)."""
        P = pend.shape[0]
        normals = npr.randn(P, self.pending_samples) if P else None
        st = self.backend.grid_state(self.covar, [(self.mean, self.noise, self.amp2, self.ls)], comp, pend, vals,
                                     normals)
        return self.backend.ei_matrix(st, cand)[:, 0]

    def optimize_hypers(self, comp, vals):
        """GPEI:348-361: a fresh gp.GP of the same kernel, ML-II from its own start values; the result replaces ours."""
        self.mean, self.noise, self.amp2, self.ls = self.backend.optimize_hypers(self.covar, comp, vals)

    # ------------------------------------------------------------------ sampling (GPEI:268-346)
    def sample_hypers(self, comp, vals):
        if self._ll is None:
            self._ll = self.backend.loglik(self.covar, comp, vals)
        if self.noiseless:
            self.noise = 1e-3
            self._sample_noiseless(comp, vals)
        else:
            self._sample_noisy(comp, vals)
        self._sample_ls(comp, vals)

    def _sample_ls(self, comp, vals):
        mean, noise, amp2, max_ls = self.mean, self.noise, self.amp2, self.max_ls

        def hypers_of(ls):
            if np.any(ls < 0) or np.any(ls > max_ls):
                return None
            return (mean, noise, amp2, ls), ()
        self.ls = util.slice_sample(self.ls, util.make_logprob(self._ll, hypers_of), compwise=True)

    def _sample_noisy(self, comp, vals):
        vmax, vmin, ls = np.max(vals), np.min(vals), self.ls

        def hypers_of(hypers):
            mean, amp2, noise = hypers[0], hypers[1], hypers[2]
            if mean > vmax or mean < vmin:
                return None
            if amp2 < 0 or noise < 0:
                return None
            return (mean, noise, amp2, ls), (
                np.log(np.log(1 + (self.noise_scale / noise) ** 2)),
                -0.5 * (np.log(amp2) / self.amp2_scale) ** 2)              # log(amp2): GPEI:312
        hypers = util.slice_sample(np.array([self.mean, self.amp2, self.noise]),
                                   util.make_logprob(self._ll, hypers_of), compwise=False)
        self.mean, self.amp2, self.noise = hypers[0], hypers[1], hypers[2]

    def _sample_noiseless(self, comp, vals):
        vmax, vmin, ls = np.max(vals), np.min(vals), self.ls

        def hypers_of(hypers):
            mean, amp2 = hypers[0], hypers[1]
            if mean > vmax or mean < vmin:
                return None
            if amp2 < 0:
                return None
            return (mean, 1e-3, amp2, ls), (-0.5 * (np.log(amp2) / self.amp2_scale) ** 2,)
        hypers = util.slice_sample(np.array([self.mean, self.amp2, self.noise]),
                                   util.make_logprob(self._ll, hypers_of), compwise=False)
        self.mean, self.amp2, self.noise = hypers[0], hypers[1], 1e-3
