"""Device engine for the GP-EI hot path: owns HBM buffers (torch tensors) and drives the C-ABI kernels.

Mirrors, per batch of hyper-samples, what the reference does one sample at a time in
``GPEIOptChooser.compute_ei`` (chooser/GPEIOptChooser.py:527-619) and
``GPEIperSecChooser.compute_ei_per_s`` (chooser/GPEIperSecChooser.py:437-548):

    cov_build -> potrf -> chol_solve (alpha) -> predict (mu, var) -> ei_sweep

HBM layout (T = float32 on the production path, float64 for log-likelihoods / logic tests):
    X        [N][D]            observed (or observed+pending) inputs, row-major
    C        [M][D]            candidates
    hypers   inv_ls[S][D], amp2[S], noise[S], mean[S]
    factor   [S][Npad][Npad]   K, overwritten by its lower Cholesky factor; Npad = ceil128(N)
    winv     [S][Npad/NB][NB][NB]  inverses of the diagonal blocks of L
    alpha    [S][F][Npad]
    mu       [S][F][ldm], var [S][ldm], ei [S][ldm], ei_sum [ldm];  ldm = ceil128(M)

torch provides the allocator and the stream only.  Nothing here falls back to the CPU.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import KINDS, check, fn, ptr


def _ceil(x, m):
    return ((x + m - 1) // m) * m


class HyperBatch(object):
    """S hyper-samples (the reference's list of (mean, noise, amp2, ls) tuples, OPT:628) on the device."""

    def __init__(self, hyper_samples, kind, device, dtype):
        S = len(hyper_samples)
        D = len(np.atleast_1d(hyper_samples[0][3]))
        mean = np.array([h[0] for h in hyper_samples], dtype=np.float64)
        noise = np.array([h[1] for h in hyper_samples], dtype=np.float64)
        amp2 = np.array([h[2] for h in hyper_samples], dtype=np.float64)
        ls = np.vstack([np.atleast_1d(h[3]) for h in hyper_samples]).astype(np.float64)
        if kind == "SE":            # gp.SE overwrites ls with ones (gp.py:88)
            ls = np.ones_like(ls)
        host = np.concatenate([mean, noise, amp2, (1.0 / ls).ravel()])
        dev = torch.from_numpy(host).to(device=device, dtype=dtype, non_blocking=False)
        self.S, self.D = S, D
        self.mean, self.noise, self.amp2 = dev[:S], dev[S:2 * S], dev[2 * S:3 * S]
        self.inv_ls = dev[3 * S:].view(S, D)
        self.host_mean, self.host_noise, self.host_amp2 = mean, noise, amp2


class Factor(object):
    """Batched Cholesky factors of K_s = amp2_s (k + 1e-6 I) + noise_s I for S hyper-samples."""

    def __init__(self, eng, kind, X, hb):
        self.eng, self.kind, self.hb = eng, kind, hb
        self.X = X
        self.N, self.D = X.shape
        self.Npad = _ceil(self.N, 128)
        S, dt, dev = hb.S, eng.dtype, eng.device
        NB = eng.NB
        self.L = torch.empty((S, self.Npad, self.Npad), dtype=dt, device=dev)
        self.winv = torch.empty((S, self.Npad // NB, NB, NB), dtype=dt, device=dev)
        self.info = torch.zeros((S,), dtype=torch.int32, device=dev)
        st = eng.stream()
        check(fn("smk_cov_build", dt)(KINDS[kind], self.N, self.N, self.D, S, ptr(X), None, ptr(hb.inv_ls),
                                      ptr(hb.amp2), ptr(hb.noise), ptr(self.L), self.Npad, st), "cov_build")
        check(fn("smk_potrf_lower_batched", dt)(self.Npad, S, ptr(self.L), ptr(self.winv), ptr(self.info), st),
              "potrf")

    def check_pd(self):
        """The reference lets spla.cholesky raise LinAlgError (SURVEY 8b 'Errors'); so do we."""
        info = self.info.cpu().numpy()
        if np.any(info != 0):
            s = int(np.nonzero(info)[0][0])
            raise np.linalg.LinAlgError("%d-th leading minor of the array is not positive definite "
                                        "(hyper-sample %d)" % (int(info[s]), s))

    def solve(self, y, F=1, y_stride=0, ldy=None, n_lead=None, want_alpha=True, want_logdet=False,
              want_quad=False, subtract_mean=True):
        """alpha = K^-1 (y - mean) for F right-hand sides; optionally sum log diag(L) and the quadratic form.

        ``n_lead`` < N solves against the leading n_lead x n_lead block of the factor (OPT:574)."""
        eng, hb = self.eng, self.hb
        S, dt, dev = hb.S, eng.dtype, eng.device
        n = self.N if n_lead is None else n_lead
        alpha = torch.empty((S, F, self.Npad), dtype=dt, device=dev) if want_alpha else None
        sld = torch.empty((S,), dtype=dt, device=dev) if want_logdet else None
        quad = torch.empty((S, F), dtype=dt, device=dev) if want_quad else None
        check(fn("smk_chol_solve", dt)(n, self.Npad, S, F, ptr(self.L), ptr(self.winv), ptr(y), y_stride,
                                       ldy if ldy is not None else n, ptr(hb.mean) if subtract_mean else None,
                                       ptr(alpha), ptr(sld), ptr(quad), eng.stream()), "chol_solve")
        return alpha, sld, quad


class GPEIEngine(object):
    """One engine per process / GPU.  ``dtype`` float32 is the product path."""

    def __init__(self, device=None, dtype=torch.float32):
        if not torch.cuda.is_available():
            raise _lib.SmkError("spearmint_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _lib.lib()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        self.dtype = dtype
        self.esize = 8 if dtype == torch.float64 else 4
        self.NB = _lib.lib().smk_block(self.esize)
        self._ws = None
        self.last = {}
        self.timers = None   # set to {} to record (start, end) CUDA events per stage on the launch stream

    def _t0(self):
        if self.timers is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream(self.device))
        return e

    def _t1(self, name, e0):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record(torch.cuda.current_stream(self.device))
        self.timers.setdefault(name, []).append((e0, e1))

    def stage_ms(self):
        """Sum of recorded stage durations in ms (call after a synchronize)."""
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in (self.timers or {}).items()}

    # ------------------------------------------------------------------ plumbing
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def to_dev(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        return torch.from_numpy(a).to(device=self.device, dtype=self.dtype)

    def workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
        return self._ws

    def max_samples_per_chunk(self, Npad, ldm, F=1):
        """How many hyper-samples fit at once (factor + winv + alpha + mu/var/ei) in ~70% of free HBM."""
        free, _ = torch.cuda.mem_get_info(self.device)
        free += torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
        per = self.esize * (Npad * Npad + Npad * self.NB + F * Npad + (F + 3) * ldm)
        fixed = _lib.lib().smk_predict_workspace_bytes(self.esize, Npad)
        return max(1, int((0.7 * free - fixed) // per))

    # ------------------------------------------------------------------ building blocks
    def hypers(self, hyper_samples, kind):
        return HyperBatch(hyper_samples, kind, self.device, self.dtype)

    def factor(self, kind, X, hb):
        return Factor(self, kind, X, hb)

    def cov(self, kind, hb, X, Y=None):
        """Batched chooser.cov (OPT:207-212): returns [S][N][N] (self, jitter included, no noise) or [S][N][M]."""
        N, D = X.shape
        dt = self.dtype
        if Y is None:
            out = torch.empty((hb.S, N, N), dtype=dt, device=self.device)
            check(fn("smk_cov_build", dt)(KINDS[kind], N, N, D, hb.S, ptr(X), None, ptr(hb.inv_ls), ptr(hb.amp2),
                                          None, ptr(out), N, self.stream()), "cov_build")
        else:
            M = Y.shape[0]
            out = torch.empty((hb.S, N, M), dtype=dt, device=self.device)
            check(fn("smk_cov_build", dt)(KINDS[kind], N, M, D, hb.S, ptr(X), ptr(Y), ptr(hb.inv_ls),
                                          ptr(hb.amp2), None, ptr(out), M, self.stream()), "cov_build")
        return out

    def predict(self, kind, fac, C_dev, alpha):
        """Predictive mean / variance at the candidates for every sample of the factor batch."""
        hb, dt = fac.hb, self.dtype
        M = C_dev.shape[0]
        ldm = _ceil(M, 128)
        mu = torch.empty((hb.S, ldm), dtype=dt, device=self.device)
        var = torch.empty((hb.S, ldm), dtype=dt, device=self.device)
        nb = _lib.lib().smk_predict_workspace_bytes(self.esize, fac.Npad)
        ws = self.workspace(nb)
        check(fn("smk_predict", dt)(KINDS[kind], fac.N, fac.Npad, M, fac.D, hb.S, ptr(fac.X), ptr(C_dev),
                                    ptr(hb.inv_ls), ptr(hb.amp2), ptr(hb.mean), ptr(fac.L), ptr(fac.winv),
                                    ptr(alpha), ptr(mu), ptr(var), ldm, ptr(ws), nb, self.stream()), "predict")
        return mu, var, ldm

    def cross_mean(self, kind, fac, C_dev, alpha, F):
        hb, dt = fac.hb, self.dtype
        M = C_dev.shape[0]
        ldm = _ceil(M, 128)
        mu = torch.empty((hb.S, F, ldm), dtype=dt, device=self.device)
        check(fn("smk_cross_mean", dt)(KINDS[kind], fac.N, fac.Npad, M, fac.D, hb.S, F, ptr(fac.X), ptr(C_dev),
                                       ptr(hb.inv_ls), ptr(hb.amp2), ptr(hb.mean), ptr(alpha), ptr(mu), ldm,
                                       self.stream()), "cross_mean")
        return mu

    def ei_sweep(self, M, S, F, mu, var, ldm, best, log_time=None, want_ei=True, ei_sum=None):
        dt = self.dtype
        ei = torch.empty((S, ldm), dtype=dt, device=self.device) if want_ei else None
        if ei_sum is None:
            ei_sum = torch.zeros((ldm,), dtype=dt, device=self.device)
        check(fn("smk_ei_sweep", dt)(M, S, F, ptr(mu), ptr(var), ldm, ptr(best), ptr(log_time), ptr(ei),
                                     ptr(ei_sum), self.stream()), "ei_sweep")
        return ei, ei_sum

    def topk(self, score, M, k):
        """Indices of the k largest scores, ascending (argsort(score)[-k:], OPT:270; [-1] is the argmax, OPT:294)."""
        dt = self.dtype
        nb = _lib.lib().smk_topk_workspace_bytes(M, k)
        ws = torch.empty((nb,), dtype=torch.uint8, device=self.device)
        idx = torch.empty((k,), dtype=torch.int32, device=self.device)
        val = torch.empty((k,), dtype=dt, device=self.device)
        check(fn("smk_topk", dt)(M, k, ptr(score), ptr(idx), ptr(val), ptr(ws), nb, self.stream()), "topk")
        return idx, val

    # ------------------------------------------------------------------ whole path
    def ei_over_hypers_device(self, kind, hyper_samples, comp, pend, cand, vals, normals=None,
                              time_hyper_samples=None, durs_log=None, want_matrix=True, inputs_on_device=None):
        """Runs the batched path; returns (ei [S][ldm] or None, ei_sum [ldm], M) as device tensors.

        ``normals`` (P,F): the fantasy standard normals the reference draws on the host (OPT:588-589).
        ``time_hyper_samples`` + ``durs_log``: EI per second (PSEC:437-548).
        ``inputs_on_device``: optional dict(X=, C=, y=) of resident tensors (bench 'value' leg)."""
        P = 0 if pend is None else int(pend.shape[0])
        S = len(hyper_samples)
        if inputs_on_device is not None:
            Xo, Cd, yd = inputs_on_device["X"], inputs_on_device["C"], inputs_on_device["y"]
            best_val = inputs_on_device["best"]
        else:
            Xo, Cd, yd = self.to_dev(comp), self.to_dev(cand), self.to_dev(vals)
            best_val = float(np.min(vals))
        N, D = Xo.shape
        M = Cd.shape[0]
        ldm = _ceil(M, 128)
        Fn = 1 if P == 0 else int(normals.shape[1])
        chunk = self.max_samples_per_chunk(_ceil(N + P, 128), ldm, Fn)
        ei_sum = torch.zeros((ldm,), dtype=self.dtype, device=self.device)
        ei_all = torch.empty((S, ldm), dtype=self.dtype, device=self.device) if want_matrix else None
        for s0 in range(0, S, chunk):
            hs = hyper_samples[s0:s0 + chunk]
            if inputs_on_device is not None and inputs_on_device.get("hb") is not None and chunk >= S:
                hb = inputs_on_device["hb"]
            else:
                hb = self.hypers(hs, kind)
            log_time = None
            if time_hyper_samples is not None:
                thb = self.hypers(time_hyper_samples[s0:s0 + chunk], kind)
                tfac = self.factor(kind, Xo, thb)
                tfac.check_pd()
                ta, _, _ = tfac.solve(self.to_dev(durs_log), F=1)
                log_time = self.cross_mean(kind, tfac, Cd, ta, 1).view(thb.S, ldm)
                del tfac
            if P == 0:
                t = self._t0()
                fac = self.factor(kind, Xo, hb)
                self._t1("cov_potrf", t)
                t = self._t0()
                alpha, _, _ = fac.solve(yd, F=1)
                self._t1("chol_solve", t)
                t = self._t0()
                mu, var, _ = self.predict(kind, fac, Cd, alpha)
                self._t1("predict", t)
                best = torch.full((hb.S, 1), best_val, dtype=self.dtype, device=self.device)
                t = self._t0()
                ei, _ = self.ei_sweep(M, hb.S, 1, mu, var, ldm, best, log_time, want_matrix, ei_sum)
                self._t1("ei_sweep", t)
                fac.check_pd()   # one host sync per chunk, after everything is queued
            else:
                ei = self._pending_chunk(kind, hb, Xo, self.to_dev(pend), Cd, yd, np.asarray(vals, float),
                                         np.asarray(normals, float), log_time, M, ldm, want_matrix, ei_sum)
            if want_matrix:
                ei_all[s0:s0 + hb.S] = ei
            self.last = dict(N=N, M=M, D=D, S=S, P=P, chunk=chunk)
        return ei_all, ei_sum, M

    def _pending_chunk(self, kind, hb, Xo, Pd, Cd, yd, vals, normals, log_time, M, ldm, want_matrix, ei_sum):
        """Pending-fantasy branch (OPT:558-619) for one chunk of hyper-samples.

        The joint (N+P) factor, all big solves and the candidate sweep run on the device; the P x P
        conditional of the pending points (P <= max_concurrent, a handful) and the fantasy draw use the
        host in float64 with the caller's normals, so the host RNG order is the reference's."""
        N, P, F = Xo.shape[0], Pd.shape[0], normals.shape[1]
        S, dt = hb.S, self.dtype
        Xj = torch.cat([Xo, Pd], dim=0).contiguous()
        fac = self.factor(kind, Xj, hb)
        fac.check_pd()
        # alpha of the observed-only system from the leading N x N block of the joint factor (OPT:574-577)
        a_obs, _, _ = fac.solve(yd, F=1, n_lead=N)
        # pend_m = pend_cross' alpha + mean (OPT:581): cross mean of the observed set at the P pending points
        ofac = _LeadingView(fac, Xo, N)
        pend_m = self.cross_mean(kind, ofac, Pd, a_obs, 1)[:, 0, :P].double().cpu().numpy()          # (S,P)
        # pend_K = Schur complement - noise I, from the trailing P x P block of the joint factor (OPT:582)
        Lpp = fac.L[:, N:N + P, N:N + P].double().cpu().numpy()
        fant = np.empty((S, F, N + P))
        bests = np.empty((S, F))
        for s in range(S):
            Lp = np.tril(Lpp[s])
            pend_K = Lp.dot(Lp.T) - hb.host_noise[s] * np.eye(P)
            pend_chol = np.linalg.cholesky(pend_K)                                # LinAlgError like OPT:585
            pf = pend_chol.dot(normals) + pend_m[s][:, None]                      # (P,F)  OPT:589
            fant[s, :, :N] = vals[None, :]
            fant[s, :, N:] = pf.T
            bests[s] = np.minimum(vals.min(), pf.min(axis=0))                     # OPT:597
        fant_d = self.to_dev(fant)                                                # [S][F][N+P]
        alpha_f, _, _ = fac.solve(fant_d, F=F, y_stride=F * (N + P), ldy=N + P)   # OPT:603
        zero_alpha = torch.zeros((S, fac.Npad), dtype=dt, device=self.device)
        _, var, _ = self.predict(kind, fac, Cd, zero_alpha)                        # OPT:605, 610
        mu = self.cross_mean(kind, fac, Cd, alpha_f, F)                            # OPT:609
        ei, _ = self.ei_sweep(M, S, F, mu, var, ldm, self.to_dev(bests), log_time, want_matrix, ei_sum)
        return ei

    def ei_over_hypers(self, kind, hyper_samples, comp, pend, cand, vals, normals=None,
                       time_hyper_samples=None, durs_log=None):
        """Host-facing: numpy in, (M,S) float64 numpy out -- the reference's ei_over_hypers contract (OPT:331-341)."""
        ei, _, M = self.ei_over_hypers_device(kind, hyper_samples, comp, pend, cand, vals, normals,
                                              time_hyper_samples, durs_log, want_matrix=True)
        return ei[:, :M].t().contiguous().double().cpu().numpy()


class _LeadingView(object):
    """A Factor-like view exposing only the first N rows of a joint factor's inputs (for cross_mean)."""

    def __init__(self, fac, X, N):
        self.hb, self.X, self.N, self.D, self.Npad = fac.hb, X, N, fac.D, fac.Npad
        self.L, self.winv = fac.L, fac.winv
