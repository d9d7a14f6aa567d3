"""Pins oracle/gp_oracle.py (the numpy restatement) to golden vectors produced by the real reference
(tests/golden/make_golden.py, executed through oracle/ref_shim.py).  CPU only."""
import numpy as np
import pytest

from oracle import gp_oracle as O
from tests.helpers import OPT_CASES, PSEC_CASES, hypers, load, sets

RTOL = 1e-9


@pytest.mark.parametrize("D", [1, 3, 8])
def test_kernels_match_reference(D):
    g = load("kernels")
    x1, x2, ls = g["D%d_x1" % D], g["D%d_x2" % D], g["D%d_ls" % D]
    np.testing.assert_allclose(O.dist2(ls, x1), g["D%d_dist2_self" % D], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(O.dist2(ls, x1, x2), g["D%d_dist2_cross" % D], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(O.grad_dist2(ls, x1, x2), g["D%d_grad_dist2" % D], rtol=1e-12, atol=1e-14)
    for k in O.KINDS:
        np.testing.assert_allclose(O.kernel(k, ls, x1), g["D%d_%s_self" % (D, k)], rtol=1e-12)
        np.testing.assert_allclose(O.kernel(k, ls, x1, x2), g["D%d_%s_cross" % (D, k)], rtol=1e-12)
        np.testing.assert_allclose(O.grad_kernel(k, ls, x1, x2), g["D%d_%s_grad" % (D, k)], rtol=1e-11, atol=1e-14)


@pytest.mark.parametrize("name", OPT_CASES)
def test_ei_over_hypers_matches_reference(name):
    g = load(name)
    comp, pend, cand, vals = sets(g)
    ei = O.ei_over_hypers(str(g["kind"]), hypers(g), comp, pend, cand, vals, g["normals"])
    scale = np.abs(g["overall_ei"]).max()
    np.testing.assert_allclose(ei, g["overall_ei"], rtol=RTOL, atol=1e-12 * scale)
    assert O.select(ei) == int(np.argmax(g["overall_ei"].mean(axis=1)))


@pytest.mark.parametrize("name", [c for c in OPT_CASES if c != "opt_d3_se"])
def test_grad_optimize_ei_matches_reference(name):
    g = load(name)
    comp, pend, cand, vals = sets(g)
    for x, f_ref, g_ref in zip(g["grad_pts"], g["grad_f"], g["grad_g"]):
        f, gr = O.grad_optimize_ei_over_hypers(str(g["kind"]), hypers(g), x, comp, pend, vals, g["normals"])
        np.testing.assert_allclose(np.ravel(f)[0], f_ref, rtol=1e-8, atol=1e-14)
        np.testing.assert_allclose(gr, g_ref, rtol=1e-7, atol=1e-12 * max(1.0, np.abs(g_ref).max()))


def test_reference_gradient_is_half_of_true_gradient():
    """the reference's analytic gradient is 0.5 x the finite-difference one."""
    g = load("opt_d8_m52")
    comp, pend, cand, vals = sets(g)
    hs = hypers(g)
    x = g["grad_pts"][0]
    f0, gr = O.grad_optimize_ei_over_hypers("Matern52", hs, x, comp, pend, vals)
    fd = np.zeros_like(x)
    for d in range(x.size):
        e = np.zeros_like(x)
        e[d] = 1e-6
        fp, _ = O.grad_optimize_ei_over_hypers("Matern52", hs, x + e, comp, pend, vals)
        fm, _ = O.grad_optimize_ei_over_hypers("Matern52", hs, x - e, comp, pend, vals)
        fd[d] = (fp - fm) / 2e-6
    np.testing.assert_allclose(gr / fd, 0.5, rtol=1e-3)


def test_logprobs_match_reference():
    g = load("logprob_d6")
    comp, vals, kind = g["comp"], g["vals"], str(g["kind"])
    for i in range(len(g["noisy_joint_lp"])):
        lp = O.logprob_noisy(kind, g["noisy_joint_x"][i], g["noisy_joint_ls"][i], comp, vals)
        np.testing.assert_allclose(lp, g["noisy_joint_lp"][i], rtol=1e-10)
    for i in range(len(g["noiseless_joint_lp"])):
        lp = O.logprob_noiseless(kind, g["noiseless_joint_x"][i], g["noiseless_joint_ls"][i], comp, vals)
        np.testing.assert_allclose(lp, g["noiseless_joint_lp"][i], rtol=1e-10)
    for mode in ("noisy", "noiseless"):
        for i in range(len(g[mode + "_ls_lp"])):
            mean, noise, amp2 = g[mode + "_ls_state"][i]
            lp = O.logprob_ls(kind, g[mode + "_ls_x"][i], mean, noise, amp2, comp, vals)
            np.testing.assert_allclose(lp, g[mode + "_ls_lp"][i], rtol=1e-10)


@pytest.mark.parametrize("name", PSEC_CASES)
def test_ei_per_second_matches_reference(name):
    g = load(name)
    comp, pend, cand, vals = sets(g)
    durs = np.log(g["durations"][g["complete"]])
    hs, ths = hypers(g), hypers(g, "ths")
    S = int(g["S"])
    if pend.shape[0] == 0:
        for s in range(S):
            e = O.compute_ei_per_s(str(g["kind"]), hs[s], ths[s], comp, pend, cand, vals, durs)
            np.testing.assert_allclose(e, g["per_sample_ei_per_s"][:, s], rtol=RTOL, atol=1e-14)
        # PSEC:302 early return -> only column 0 filled, from the OLDEST time sample (PSEC:288)
        ov = O.ei_over_hypers_per_s(str(g["kind"]), hs, ths, comp, pend, cand, vals, durs, mcmc_iters=S)
        np.testing.assert_allclose(ov, g["overall_ei"], rtol=RTOL, atol=1e-14)
    assert np.all(g["overall_ei"][:, 1:] == 0.0)
    assert int(g["n_time_samples"]) > S  # stale burn-in time samples are never cleared (PSEC:199)


def test_sobol_oracle_matches_reference_generator():
    """Row f4: the numpy restatement of sobol_lib.i4_sobol_generate against golden points of the real generator."""
    from oracle import sobol_oracle as SO
    g = load("sobol")
    for i in range(7):
        m, n, skip = (int(x) for x in g["case%d_args" % i])
        assert np.array_equal(SO.i4_sobol_generate(m, n, skip), g["case%d_pts" % i]), (m, n, skip)


MLL_CASES = ["mll_d3_m52", "mll_d5_ardse", "mll_d2_m32"]


@pytest.mark.parametrize("name", MLL_CASES)
def test_mll_oracle_matches_reference_optimize_hypers(name):
    """Row f3: GP.optimize_hypers' objective, its 'gradient' (incl. the reference's length-scale expression) and the
    optimum, against values frozen from the real gp.py (tests/golden/make_golden_mll.py)."""
    g = load(name)
    comp, vals, kind = g["grid"][g["complete"]], g["values"][g["complete"]], str(g["kind"])
    mean = np.mean(vals)
    for pt, f_ref, g_ref in zip(g["pts"], g["f"], g["g"]):
        f, gr = O.mll_value_grad(kind, pt, comp, vals, mean)
        np.testing.assert_allclose(f, f_ref, rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(gr, g_ref, rtol=1e-8, atol=1e-9)
    m, noise, amp2, ls = O.gp_optimize_hypers(kind, comp, vals)
    np.testing.assert_allclose([m, noise, amp2], [g["opt_mean"], g["opt_noise"], g["opt_amp2"]], rtol=1e-6)
    np.testing.assert_allclose(ls, g["opt_ls"], rtol=1e-6)
