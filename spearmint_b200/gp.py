"""Drop-in for the reference's ``gp.GP`` class (spearmint/spearmint/gp.py:134-292): ML-II hyper-parameters.

    g = GP("Matern52"); g.real_init(D, vals); g.optimize_hypers(comp, vals)   ->  g.mean, g.amp2, g.noise, g.ls
    g.logprob(comp, vals)                                                     ->  GP log marginal likelihood

``optimize_hypers`` keeps the reference's host logic -- scipy L-BFGS-B over [log amp2, log noise, log ls] from
(std(vals), 1e-3, ones) inside [-10, 10] x [-10, 10] x [-10, 5]^D, the memoised factor, the jittered Cholesky (1e-8
grown by 1.1x until the matrix factors, GP:187-203) -- and takes every number from the GPU in float64:
covariance build + Cholesky + solves (the kernels of the log-likelihood path) and the traces of ``grad_nlogprob``
(csrc/grad.cu: smk_mll_grad_terms), including the reference's own length-scale expression (GP:258-259), which is not
the derivative of the likelihood but is what the reference optimises with.  No CPU fallback.
"""
import numpy as np
import scipy.optimize as spo

COVARS = ("SE", "ARDSE", "Matern32", "Matern52")


class GP(object):

    def __init__(self, covar="Matern52", mcmc_iters=10, noiseless=False, device=None, engine=None):
        if covar not in COVARS:
            raise KeyError(covar)                     # globals()[covar] in the reference (GP:136)
        self.covar = covar
        self.mcmc_iters = int(mcmc_iters)
        self.D = -1
        self.hyper_iters = 1
        self.noiseless = bool(int(noiseless))
        self.hyper_samples = []
        self.noise_scale = 0.1
        self.amp2_scale = 1
        self.max_ls = 2
        self._device, self._eng = device, engine
        self.stats = {}

    @property
    def eng(self):
        if self._eng is None:
            import torch
            from .engine import GPEIEngine
            self._eng = GPEIEngine(device=self._device, dtype=torch.float64)
        return self._eng

    def real_init(self, dims, values):
        self.D = dims
        self.ls = np.ones(self.D)
        self.amp2 = np.std(values)
        self.noise = 1e-3
        self.mean = np.mean(values)

    def cov(self, x1, x2=None):
        """amp2 (k + 1e-6 I) or amp2 k(x1, x2)  (GP:162-167)."""
        eng = self.eng
        hb = eng.hypers([(self.mean, self.noise, self.amp2, self.ls)], self.covar)
        out = eng.cov(self.covar, hb, eng.to_dev(x1), None if x2 is None else eng.to_dev(x2))
        return out[0].cpu().numpy()

    def logprob(self, comp, vals):
        """-sum log diag chol - 0.5 (y - mean)' K^-1 (y - mean)  (GP:169-179)."""
        return float(self.eng.loglik(self.covar, comp, vals)(self.mean, self.noise, self.amp2, self.ls))

    # ------------------------------------------------------------------ GP.optimize_hypers (GP:181-292)
    def _factor(self, comp_dev, mean, amp2, noise, ls):
        """jitter_chol (GP:187-203): factor of covmat + jitter I with jitter 1e-8, 1.1e-8, ... (None past 1e5)."""
        eng = self.eng
        jitter = 1e-8
        while True:
            if jitter > 100000:
                return None
            hb = eng.hypers([(mean, noise + jitter, amp2, ls)], self.covar)
            fac = eng.factor(self.covar, comp_dev, hb)
            if int(fac.info.cpu()[0]) == 0:
                return fac
            jitter = jitter * 1.1

    def value_grad(self, hypers, comp_dev, y_dev, eye_dev, mean):
        """(nlogprob, grad_nlogprob) at log-hypers (GP:222-264)."""
        import torch
        from . import _lib
        from ._lib import KINDS, check, fn, ptr
        eng = self.eng
        amp2, noise, ls = float(np.exp(hypers[0])), float(np.exp(hypers[1])), np.exp(hypers[2:])
        N, D = comp_dev.shape
        fac = self._factor(comp_dev, mean, amp2, noise, ls)
        if fac is None:        # past jitter 1e5 the reference continues with the factor of the identity (GP:191-193)
            diffs = y_dev - mean
            f = 0.5 * float((diffs * diffs).sum())
            alpha_t, kinv_t, lda, ldk = diffs.view(1, N).contiguous(), eye_dev.view(1, N, N), N, N
        else:
            alpha, sld_t, quad = fac.solve(y_dev, F=1, want_logdet=True, want_quad=True)
            kinv_t, _, _ = fac.solve(eye_dev, F=N, y_stride=0, ldy=N, subtract_mean=False)      # [1][N][Npad] = K^-1
            f = float(sld_t[0]) + 0.5 * float(quad[0, 0])
            alpha_t, lda, ldk = alpha.view(1, fac.Npad), fac.Npad, fac.Npad
        hb = eng.hypers([(mean, noise, amp2, ls)], self.covar)
        out = torch.empty((1, D + 2), dtype=torch.float64, device=eng.device)
        check(fn("smk_mll_grad_terms", eng.dtype)(KINDS[self.covar], N, D, 1, ptr(comp_dev), ptr(hb.inv_ls), ptr(alpha_t), lda,
                                                  ptr(kinv_t), ldk, ptr(out), eng.stream()), "mll_grad_terms")
        t = out.cpu().numpy()[0]
        grad = np.empty(D + 2)
        grad[0] = 0.5 * t[0] * amp2
        grad[1] = 0.5 * t[1] * noise
        grad[2:] = -amp2 * t[2:]
        self.stats["evals"] = self.stats.get("evals", 0) + 1
        return f, -grad

    def optimize_hypers(self, comp, vals):
        eng = self.eng
        import torch
        self.mean = np.mean(vals)
        comp_dev, y_dev = eng.to_dev(comp), eng.to_dev(vals)
        N = comp.shape[0]
        eye_dev = torch.eye(N, dtype=eng.dtype, device=eng.device)
        memo = {}

        def vg(h):                       # nlogprob and grad_nlogprob share one factorisation per point (memoize, GP:205-220)
            k = np.asarray(h, dtype=np.float64).tobytes()
            if memo.get("k") != k:
                memo["k"], memo["v"] = k, self.value_grad(np.asarray(h, dtype=np.float64), comp_dev, y_dev, eye_dev, self.mean)
            return memo["v"]

        self.ls = np.ones(self.D)
        self.amp2 = np.std(vals)
        self.noise = 1e-3
        hypers = np.zeros(self.ls.shape[0] + 2)
        hypers[0] = np.log(self.amp2)
        hypers[1] = np.log(self.noise)
        hypers[2:] = np.log(self.ls)
        b = [(-10, 10), (-10, 10)] + [(-10, 5)] * comp.shape[1]
        self.stats["evals"] = 0
        res = spo.fmin_l_bfgs_b(lambda h: vg(h)[0], hypers, lambda h: vg(h)[1], args=(), bounds=b)
        hypers = res[0]
        self.amp2 = np.exp(hypers[0])
        self.noise = np.exp(hypers[1])
        self.ls = np.exp(hypers[2:])
