"""-m gpu: the drop-in chooser with the real DeviceBackend against golden next() outputs of the reference, plus the
two 'next' rows it depends on: the float64 GP log-likelihood (f2) and the cached-factor EI gradient (f1)."""
import numpy as np
import pytest

from tests.helpers import hypers, load, sets

pytestmark = pytest.mark.gpu

NEXT_CASES = ["opt_branin2d", "opt_d8_m52", "opt_d8_m52_pend", "opt_d5_ardse", "opt_d4_m32_pend", "opt_d1_m52"]


def _same_proposal(ret, g):
    """The proposed POINT must be the reference's.  A grid index and an off-grid tuple can denote the same
    coordinates (e.g. a refined point clamped onto the Sobol origin): then the reference's own choice between them
    is decided by 1e-17 BLAS round-off and either form is accepted."""
    grid = g["grid"]
    ref_pt = g["next_point"] if int(g["next_is_tuple"]) else grid[int(g["next_index"])]
    got_pt = ret[1] if isinstance(ret, tuple) else grid[ret]
    np.testing.assert_allclose(got_pt, ref_pt, rtol=0, atol=2e-4)
    if isinstance(ret, tuple):
        assert ret[0] == g["candidates"].shape[0]          # (numcand, point) protocol, OPT:296-297
    same_form = isinstance(ret, tuple) == bool(int(g["next_is_tuple"]))
    if not same_form:                                      # only legitimate when the point is also a grid candidate
        cand = grid[g["candidates"]]
        assert np.min(np.abs(cand - ref_pt).max(axis=1)) < 2e-4


@pytest.fixture(scope="module")
def backend():
    from spearmint_b200.backend import DeviceBackend
    return DeviceBackend()


def test_loglik_matches_reference_logprobs(backend):
    g = load("logprob_d6")
    comp, vals, kind = g["comp"], g["vals"], str(g["kind"])
    ll = backend.loglik(kind, comp, vals)
    for i in range(len(g["noisy_ls_lp"])):
        mean, noise, amp2 = g["noisy_ls_state"][i]
        ls = g["noisy_ls_x"][i]
        if np.any(ls < 0) or np.any(ls > 2):
            continue
        np.testing.assert_allclose(ll(mean, noise, amp2, ls), g["noisy_ls_lp"][i], rtol=1e-10, atol=1e-9)
    n = 0
    for i in range(len(g["noiseless_joint_lp"])):
        mean, amp2 = g["noiseless_joint_x"][i][:2]
        if not np.isfinite(g["noiseless_joint_lp"][i]):
            continue
        lp = ll(mean, 1e-3, amp2, g["noiseless_joint_ls"][i]) - 0.5 * np.log(np.sqrt(amp2)) ** 2
        np.testing.assert_allclose(lp, g["noiseless_joint_lp"][i], rtol=1e-10, atol=1e-9)
        n += 1
    assert n > 3 and ll.calls > 10


@pytest.mark.parametrize("name", [c for c in NEXT_CASES])
def test_refine_value_grad_matches_reference(backend, name):
    g = load(name)
    comp, pend, cand, vals = sets(g)
    ctx = backend.refine_context(str(g["kind"]), hypers(g), comp, pend, vals,
                                 g["normals"] if pend.shape[0] else None)
    for x, f_ref, g_ref in zip(g["grad_pts"], g["grad_f"], g["grad_g"]):
        f, gr = ctx.value_grad(x)
        np.testing.assert_allclose(f, f_ref, rtol=1e-7, atol=1e-12)
        np.testing.assert_allclose(gr, g_ref, rtol=1e-6, atol=1e-10 * max(1.0, np.abs(g_ref).max()))


@pytest.mark.parametrize("name", NEXT_CASES)
def test_next_matches_reference(backend, name, tmp_path):
    from spearmint_b200.chooser import GPEIOptChooserB200 as mod
    g = load(name)
    args = "covar=%s,mcmc_iters=%d,burnin=%d,noiseless=%d,use_multiprocessing=0,grid_subset=5" % (
        str(g["kind"]), int(g["S"]), int(g["burnin"]), int(g["noiseless"]))
    ch = mod.init(str(tmp_path), args)
    ch._backend = backend
    np.random.seed(int(g["seed"]))
    ret = ch.next(g["grid"], g["values"], g["durations"], g["candidates"], g["pending"], g["complete"])
    for a, b in zip(ch.hyper_samples, hypers(g)):       # float64 GPU log-likelihood -> same chain
        np.testing.assert_allclose(np.hstack(a), np.hstack(b), rtol=1e-6, atol=1e-9)
    _same_proposal(ret, g)


def test_public_ei_methods_match_reference(backend, tmp_path):
    from spearmint_b200.chooser import GPEIOptChooserB200 as mod
    g = load("opt_d8_m52_pend")
    comp, pend, cand, vals = sets(g)
    ch = mod.init(str(tmp_path), "mcmc_iters=%d,noiseless=1" % int(g["S"]))
    ch._backend = backend
    np.random.seed(3)
    ch._real_init(comp.shape[1], vals)
    # the reference captured randomstate at its first next(); the golden stores the resulting normals instead
    ch._fantasy_normals = lambda p: g["normals"] if p.shape[0] else None
    ch.hyper_samples = hypers(g)
    ei = ch.ei_over_hypers(comp, pend, cand, vals)
    ref = g["overall_ei"]
    assert np.abs(ei - ref).max() <= 5e-3 * ref.max()
    assert ch.mean == ch.hyper_samples[-1][0]
    one = ch.compute_ei(comp, pend, cand, vals)
    assert np.abs(one - ref[:, -1]).max() <= 5e-3 * ref[:, -1].max()


@pytest.mark.parametrize("name", ["psec_d4", "psec_d3_pend"])
def test_per_second_next_matches_reference(backend, name, tmp_path):
    from spearmint_b200.chooser import GPEIperSecChooserB200 as mod
    g = load(name)
    ch = mod.init(str(tmp_path), "covar=%s,mcmc_iters=%d,burnin=%d,grid_subset=4" % (
        str(g["kind"]), int(g["S"]), int(g["burnin"])))
    ch._backend = backend
    np.random.seed(int(g["seed"]))
    ret = ch.next(g["grid"], g["values"], g["durations"], g["candidates"], g["pending"], g["complete"])
    for a, b in zip(ch.hyper_samples, hypers(g)):
        np.testing.assert_allclose(np.hstack(a), np.hstack(b), rtol=1e-6, atol=1e-9)
    for a, b in zip(ch.time_hyper_samples, hypers(g, "ths")):
        np.testing.assert_allclose(np.hstack(a), np.hstack(b), rtol=1e-6, atol=1e-9)
    _same_proposal(ret, g)


def test_per_second_refine_gradient(backend):
    g = load("psec_d4")
    comp, pend, cand, vals = sets(g)
    durs = np.log(g["durations"][g["complete"]])
    S = int(g["S"])
    ctx = backend.refine_context(str(g["kind"]), hypers(g)[:S], comp, np.zeros((0, comp.shape[1])), vals, None,
                                 hypers(g, "ths")[:S], durs)
    for x, f_ref, g_ref in zip(g["grad_pts"], g["grad_f"], g["grad_g"]):
        f, gr = ctx.value_grad(x)
        np.testing.assert_allclose(f, f_ref, rtol=1e-7)
        np.testing.assert_allclose(gr, g_ref, rtol=1e-6, atol=1e-11)


def test_gpei_chooser_next_matches_reference(backend, tmp_path):
    from spearmint_b200.chooser import GPEIChooserB200 as mod
    g = load("gpei_d3")
    ch = mod.init(str(tmp_path), "mcmc_iters=4")
    ch._backend = backend
    np.random.seed(int(g["seed"]))
    ret = ch.next(g["grid"], g["values"], g["durations"], g["candidates"], g["pending"], g["complete"])
    assert ret == int(g["next_index"])
