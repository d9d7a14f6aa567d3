"""Drop-in replacement for the reference's ``chooser/GPEIperSecChooser.py`` (EI per second, two GPs).

Same plugin API, options, state pickle keys (dims, ls, amp2, noise, mean, time_ls, time_amp2, time_noise,
time_mean -- PSEC:87-95) and global-RNG order.  The reference's observable quirks are reproduced on purpose,
because they decide which point it proposes:
  * ``ei_over_hypers`` returns from INSIDE its sample loop (PSEC:302): only column 0 is ever filled;
  * ``time_hyper_samples`` is never cleared (PSEC:199), so index [i] reads the OLDEST (burn-in) time samples (PSEC:288);
  * the refinement ignores pending points (PSEC:351-435) and is serial;
  * fantasy normals come straight from the global RNG stream, with no state reset (PSEC:521);
  * max_ls = 10; objective amp2 prior log(amp2) (PSEC:614), time-GP amp2 prior log(sqrt(amp2)) (PSEC:646).
All arithmetic runs on the GPU through spearmint_b200.backend.DeviceBackend.
"""
import os
import pickle
import tempfile

import numpy as np
import numpy.random as npr
import scipy.optimize as spo

from spearmint_b200 import util
from spearmint_b200.locker import Locker, log

COVARS = ("SE", "ARDSE", "Matern32", "Matern52")


def init(expt_dir, arg_string):
    args = util.unpack_args(arg_string)
    return GPEIperSecChooserB200(expt_dir, **args)


class GPEIperSecChooserB200(object):

    def __init__(self, expt_dir, covar="Matern52", mcmc_iters=10, pending_samples=100, noiseless=False, burnin=100,
                 grid_subset=20, device=None, refine_dtype="float64", state_name=None, backend=None):
        if covar not in COVARS:
            raise AttributeError("module 'spearmint.gp' has no attribute '%s'" % covar)
        self.covar = covar
        self.locker = Locker()
        name = state_name if state_name else self.__module__
        self.state_pkl = os.path.join(expt_dir, name + ".pkl")
        self.stats_file = os.path.join(expt_dir, name + "_hyperparameters.txt")
        self.mcmc_iters = int(mcmc_iters)
        self.burnin = int(burnin)
        self.needs_burnin = True
        self.pending_samples = int(pending_samples)
        self.D = -1
        self.hyper_iters = 1
        self.grid_subset = int(grid_subset)
        self.noiseless = bool(int(noiseless))
        self.hyper_samples = []
        self.time_hyper_samples = []
        self.noise_scale, self.amp2_scale, self.max_ls = 0.1, 1, 10
        self.time_noise_scale, self.time_amp2_scale, self.time_max_ls = 0.1, 1, 10
        self._device, self._refine_dtype = device, refine_dtype
        self._backend = backend
        self._ll_obj = self._ll_time = None

    @property
    def backend(self):
        if self._backend is None:
            from spearmint_b200.backend import DeviceBackend
            self._backend = DeviceBackend(device=self._device, refine_dtype=self._refine_dtype)
        return self._backend

    # ------------------------------------------------------------------ state (PSEC:82-143)
    def dump_hypers(self):
        self.locker.lock_wait(self.state_pkl)
        fh = tempfile.NamedTemporaryFile(mode="wb", delete=False)
        pickle.dump({"dims": self.D, "ls": self.ls, "amp2": self.amp2, "noise": self.noise, "mean": self.mean,
                     "time_ls": self.time_ls, "time_amp2": self.time_amp2, "time_noise": self.time_noise,
                     "time_mean": self.time_mean}, fh, protocol=2)
        fh.close()
        os.system('mv "%s" "%s"' % (fh.name, self.state_pkl))
        self.locker.unlock(self.state_pkl)

    def _real_init(self, dims, values, durations):
        self.locker.lock_wait(self.state_pkl)
        if os.path.exists(self.state_pkl):
            with open(self.state_pkl, "rb") as fh:
                state = pickle.load(fh)
            self.D = state["dims"]
            self.ls, self.amp2, self.noise, self.mean = state["ls"], state["amp2"], state["noise"], state["mean"]
            self.time_ls, self.time_amp2 = state["time_ls"], state["time_amp2"]
            self.time_noise, self.time_mean = state["time_noise"], state["time_mean"]
        else:
            self.D = dims
            self.ls = np.ones(self.D)
            self.time_ls = np.ones(self.D)
            self.amp2 = np.std(values) + 1e-4
            self.time_amp2 = np.std(durations) + 1e-4      # std of the RAW durations (PSEC:132)
            self.noise = 1e-3
            self.time_noise = 1e-3
            self.mean = np.mean(values)
            self.time_mean = np.mean(np.log(durations))
        self.locker.unlock(self.state_pkl)

    # ------------------------------------------------------------------ plugin entry point (PSEC:155-281)
    def next(self, grid, values, durations, candidates, pending, complete):
        if complete.shape[0] < 2:
            return int(candidates[0])
        if self.D == -1:
            self._real_init(grid.shape[1], values[complete], durations[complete])

        comp = grid[complete, :]
        cand = grid[candidates, :]
        pend = grid[pending, :]
        vals = values[complete]
        durs = np.log(durations[complete])          # log domain keeps durations positive (PSEC:176)

        numcand = cand.shape[0]
        best_comp = np.argmin(vals)
        cand2 = np.vstack((np.random.randn(10, comp.shape[1]) * 0.001 + comp[best_comp, :], cand))

        if self.mcmc_iters <= 0:
            raise NotImplementedError("mcmc_iters=0 is broken in the reference (GPEIperSecChooser.py:254 reads an "
                                      "undefined overall_ei) and not provided here")

        self._ll_obj = self.backend.loglik(self.covar, comp, vals)
        self._ll_time = self.backend.loglik(self.covar, comp, durs.squeeze())
        if self.needs_burnin:
            for mcmc_iter in range(self.burnin):
                self.sample_hypers(comp, vals, durs)
                log("BURN %d/%d] mean: %.2f  amp: %.2f noise: %.4f  min_ls: %.4f  max_ls: %.4f"
                    % (mcmc_iter + 1, self.burnin, self.mean, np.sqrt(self.amp2), self.noise,
                       np.min(self.ls), np.max(self.ls)))
            self.needs_burnin = False

        self.hyper_samples = []                      # time_hyper_samples is NOT cleared (PSEC:199)
        for mcmc_iter in range(self.mcmc_iters):
            self.sample_hypers(comp, vals, durs)
            log("%d/%d] mean: %.2f  amp: %.2f  noise: %.4f min_ls: %.4f  max_ls: %.4f"
                % (mcmc_iter + 1, self.mcmc_iters, self.mean, np.sqrt(self.amp2), self.noise,
                   np.min(self.ls), np.max(self.ls)))
            log("%d/%d] time_mean: %.2fs time_amp: %.2f  time_noise: %.4f time_min_ls: %.4f  time_max_ls: %.4f"
                % (mcmc_iter + 1, self.mcmc_iters, np.exp(self.time_mean), np.sqrt(self.time_amp2),
                   np.exp(self.time_noise), np.min(self.time_ls), np.max(self.time_ls)))
        self.dump_hypers()
        self._ll_obj = self._ll_time = None

        # grid pass 1: only sample 0 contributes (PSEC:302) -> ranking by column 0
        k = min(self.grid_subset, cand2.shape[0])
        inds = self.backend.top_mean_ei(self._grid_state(comp, pend, vals, durs), cand2, k)
        cand2 = cand2[inds, :]

        b = [(0, 1)] * cand.shape[1]
        ctx = self.backend.refine_context(self.covar, self.hyper_samples[:self.mcmc_iters], comp,
                                          np.zeros((0, comp.shape[1])), vals, None,
                                          self.time_hyper_samples[:self.mcmc_iters], durs)
        for i in range(cand2.shape[0]):
            log("Optimizing candidate %d/%d" % (i + 1, cand2.shape[0]))
            ret = spo.fmin_l_bfgs_b(ctx.value_grad, cand2[i, :].flatten(), bounds=b)
            cand2[i, :] = ret[0]
        del ctx
        cand = np.vstack((cand, cand2))

        best_cand = int(self.backend.top_mean_ei(self._grid_state(comp, pend, vals, durs), cand, 1)[-1])
        self._load_pair(0)
        self.dump_hypers()
        if best_cand >= numcand:
            return (int(numcand), cand[best_cand, :])
        return int(candidates[best_cand])

    # ------------------------------------------------------------------ EI per second over hyper-samples (PSEC:284-302)
    def _load_pair(self, i):
        (self.mean, self.noise, self.amp2, self.ls) = self.hyper_samples[i]
        (self.time_mean, self.time_noise, self.time_amp2, self.time_ls) = self.time_hyper_samples[i]

    def _grid_state(self, comp, pend, vals, durs):
        normals = npr.randn(pend.shape[0], self.pending_samples) if pend.shape[0] else None    # PSEC:521
        return self.backend.grid_state(self.covar, [self.hyper_samples[0]], comp, pend, vals, normals,
                                       [self.time_hyper_samples[0]], np.asarray(durs).squeeze())

    def ei_over_hypers(self, comp, pend, cand, vals, durs):
        """(M, mcmc_iters) with ONLY column 0 filled -- the reference returns inside its loop (PSEC:302)."""
        out = np.zeros((cand.shape[0], self.mcmc_iters))
        out[:, 0] = self.backend.ei_matrix(self._grid_state(comp, pend, vals, durs), cand)[:, 0]
        self._load_pair(0)
        return out

    def compute_ei_per_s(self, comp, pend, cand, vals, durs):
        """EI / exp(predicted log-duration) under the CURRENT hyper-parameters (PSEC:437-548)."""
        normals = npr.randn(pend.shape[0], self.pending_samples) if pend.shape[0] else None
        st = self.backend.grid_state(self.covar, [(self.mean, self.noise, self.amp2, self.ls)], comp, pend, vals,
                                     normals, [(self.time_mean, self.time_noise, self.time_amp2, self.time_ls)],
                                     np.asarray(durs).squeeze())
        return self.backend.ei_matrix(st, cand)[:, 0]

    def grad_optimize_ei_over_hypers(self, cand, comp, vals, durs, compute_grad=True):
        ctx = self.backend.refine_context(self.covar, self.hyper_samples[:self.mcmc_iters], comp,
                                          np.zeros((0, comp.shape[1])), vals, None,
                                          self.time_hyper_samples[:self.mcmc_iters], durs)
        f, g = ctx.value_grad(cand)
        return (f, g) if compute_grad else f

    # ------------------------------------------------------------------ sampling (PSEC:550-681)
    def sample_hypers(self, comp, vals, durs):
        if self._ll_obj is None:
            self._ll_obj = self.backend.loglik(self.covar, comp, vals)
            self._ll_time = self.backend.loglik(self.covar, comp, np.asarray(durs).squeeze())
        if self.noiseless:
            self.noise = 1e-3
            self._sample_noiseless(comp, vals)
        else:
            self._sample_noisy(comp, vals)
        self._sample_ls(comp, vals)
        self._sample_time_noisy(comp, np.asarray(durs).squeeze())
        self._sample_time_ls(comp, np.asarray(durs).squeeze())
        self.hyper_samples.append((self.mean, self.noise, self.amp2, self.ls))
        self.time_hyper_samples.append((self.time_mean, self.time_noise, self.time_amp2, self.time_ls))

    def _sample_ls(self, comp, vals):
        mean, noise, amp2, max_ls = self.mean, self.noise, self.amp2, self.max_ls

        def hypers_of(ls):
            if np.any(ls < 0) or np.any(ls > max_ls):
                return None
            return (mean, noise, amp2, ls), ()
        self.ls = util.slice_sample(self.ls, util.make_logprob(self._ll_obj, hypers_of), compwise=True)

    def _sample_time_ls(self, comp, durs):
        mean, noise, amp2, max_ls = self.time_mean, self.time_noise, self.time_amp2, self.time_max_ls

        def hypers_of(ls):
            if np.any(ls < 0) or np.any(ls > max_ls):
                return None
            return (mean, noise, amp2, ls), ()
        self.time_ls = util.slice_sample(self.time_ls, util.make_logprob(self._ll_time, hypers_of), compwise=True)

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
                -0.5 * (np.log(amp2) / self.amp2_scale) ** 2)               # log(amp2), not log(sqrt(amp2)): PSEC:614
        hypers = util.slice_sample(np.array([self.mean, self.amp2, self.noise]),
                                   util.make_logprob(self._ll_obj, hypers_of), compwise=False)
        self.mean, self.amp2, self.noise = hypers[0], hypers[1], hypers[2]

    def _sample_time_noisy(self, comp, durs):
        vmax, vmin, ls = np.max(durs), np.min(durs), self.time_ls

        def hypers_of(hypers):
            mean, amp2, noise = hypers[0], hypers[1], hypers[2]
            if mean > vmax or mean < vmin:
                return None
            if amp2 < 0 or noise < 0:
                return None
            return (mean, noise, amp2, ls), (
                np.log(np.log(1 + (self.time_noise_scale / noise) ** 2)),
                -0.5 * (np.log(np.sqrt(amp2)) / self.time_amp2_scale) ** 2)   # PSEC:646
        hypers = util.slice_sample(np.array([self.time_mean, self.time_amp2, self.time_noise]),
                                   util.make_logprob(self._ll_time, hypers_of), compwise=False)
        self.time_mean, self.time_amp2, self.time_noise = hypers[0], hypers[1], hypers[2]

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
                                   util.make_logprob(self._ll_obj, hypers_of), compwise=False)
        self.mean, self.amp2, self.noise = hypers[0], hypers[1], 1e-3
