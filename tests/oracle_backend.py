"""TEST-ONLY stand-in for spearmint_b200.backend.DeviceBackend built on the CPU oracle, so the chooser's host logic
(RNG order, state files, return protocol, quirk reproduction) can be checked without a GPU.  Never used by the product."""
import numpy as np

from oracle import gp_oracle as O


class _State(object):
    pass


class OracleBackend(object):
    name = "oracle"

    def __init__(self, batched=False, speculate=None):
        self.loglik_calls = 0
        self.batches = 0
        self.batched = batched
        self.speculate = speculate

    def loglik(self, kind, comp, vals):
        outer = self

        def ll(mean, noise, amp2, ls):
            outer.loglik_calls += 1
            return O.gp_logprob(kind, mean, noise, amp2, np.asarray(ls, float), comp, vals)
        if not self.batched:
            return ll

        class Batched(object):          # same interface as engine.LogLik: exercises the speculative sampler path
            if outer.speculate:
                speculate = outer.speculate
            def __call__(self, mean, noise, amp2, ls):
                return ll(mean, noise, amp2, ls)

            def batch(self, hypers):
                outer.batches += 1
                out = np.empty(len(hypers))
                for i, h in enumerate(hypers):
                    try:
                        out[i] = ll(*h)
                    except np.linalg.LinAlgError:
                        out[i] = np.nan
                return out
        return Batched()

    def optimize_hypers(self, kind, comp, vals):
        return O.gp_optimize_hypers(kind, comp, vals)

    def grid_state(self, kind, hyper_samples, comp, pend, vals, normals=None, time_hyper_samples=None, durs_log=None):
        st = _State()
        st.kind, st.hs, st.comp, st.pend, st.vals, st.normals = kind, list(hyper_samples), comp, pend, vals, normals
        st.ths, st.durs = time_hyper_samples, durs_log
        return st

    def ei_matrix(self, st, cand):
        if st.ths is None and (st.normals is None or np.ndim(st.normals) == 2):
            return O.ei_over_hypers(st.kind, st.hs, st.comp, st.pend, cand, st.vals, st.normals)
        if st.ths is None:      # per-sample fantasy normals (GPEIChooser)
            out = np.zeros((cand.shape[0], len(st.hs)))
            for s, h in enumerate(st.hs):
                out[:, s] = O.compute_ei(st.kind, h, st.comp, st.pend, cand, st.vals, st.normals[s])
            return out
        out = np.zeros((cand.shape[0], len(st.hs)))
        for s, (h, th) in enumerate(zip(st.hs, st.ths)):
            out[:, s] = O.compute_ei_per_s(st.kind, h, th, st.comp, st.pend, cand, st.vals, st.durs, st.normals)
        return out

    def top_mean_ei(self, st, cand, k):
        # same tie rule as the device top-k: ascending score, ties ordered so that the LOWER index ranks higher
        # (so [-1] is numpy's first-max argmax, OPT:294)
        m = np.mean(self.ei_matrix(st, cand), axis=1)
        return np.lexsort((-np.arange(m.size), m))[-k:]

    def refine_context(self, kind, hyper_samples, comp, pend, vals, normals=None, time_hyper_samples=None,
                       durs_log=None):
        hs = list(hyper_samples)
        outer = self

        class Ctx(object):
            evals = 0

            def value_grad(self, x):
                self.evals += 1
                if time_hyper_samples is None:
                    f, g = O.grad_optimize_ei_over_hypers(kind, hs, x, comp, pend, vals, normals)
                else:
                    f, g = O.grad_optimize_ei_per_s_over_hypers(kind, hs, list(time_hyper_samples), x, comp, vals,
                                                                durs_log)
                return float(np.ravel(f)[0]), g
        return Ctx()
