"""Host helpers of the chooser plugin API: the "k=v,k=v" argument protocol and the slice sampler.

Both are part of the drop-in boundary: ``init(expt_dir, arg_string)`` hands every option over as a
string, and the hyper-parameter chain must consume the process-global numpy RNG in exactly the reference's order
(util.py:34-93 of the reference) so that, with a float64 log-likelihood, the chain reproduces the reference's draws.
The log-probability callbacks are evaluated on the GPU (engine.LogLik); the sampler control flow stays on the host.
"""
import re

import numpy as np
import numpy.random as npr


def unpack_args(arg_string):
    """"a=1,b=x" -> {"a": "1", "b": "x"}; strings of length <= 1 give {} (reference util.py:26-32)."""
    if len(arg_string) > 1:
        pairs = re.split(r"\s*,\s*", arg_string)
        return dict(re.split(r"\s*=\s*", kv) for kv in pairs)
    return {}


def _slice_along(direction, x0, logprob, sigma, step_out, max_steps_out):
    """One slice-sampling move along ``direction`` (reference util.py:36-75).  RNG call order:
    rand (interval placement), rand (slice height), then one rand per shrink proposal."""
    def lp(z):
        return logprob(direction * z + x0)

    upper = sigma * npr.rand()
    lower = upper - sigma
    height = np.log(npr.rand()) + lp(0.0)
    n_lo = n_hi = 0
    if step_out:
        while lp(lower) > height and n_lo < max_steps_out:
            n_lo += 1
            lower -= sigma
        while lp(upper) > height and n_hi < max_steps_out:
            n_hi += 1
            upper += sigma
    while True:
        z = (upper - lower) * npr.rand() + lower
        val = lp(z)
        if np.isnan(val):
            raise Exception("Slice sampler got a NaN")
        if val > height:
            return z * direction + x0
        if z < 0:
            lower = z
        elif z > 0:
            upper = z
        else:
            raise Exception("Slice sampler shrank to zero!")


def slice_sample(init_x, logprob, sigma=1.0, step_out=True, max_steps_out=1000, compwise=False, verbose=False):
    """Univariate slice sampling along a random direction, or component-wise in shuffled order
    (reference util.py:34-93).  Acceptance is a strict ``>`` on log-probabilities."""
    init_x = np.asarray(init_x, dtype=float)
    if not init_x.shape:
        init_x = np.array([init_x])
    dims = init_x.shape[0]
    if compwise:
        order = list(range(dims))
        npr.shuffle(order)
        x = init_x.copy()
        for d in order:
            e = np.zeros(dims)
            e[d] = 1.0
            x = _slice_along(e, x, logprob, sigma, step_out, max_steps_out)
        return x
    direction = npr.randn(dims)
    direction = direction / np.sqrt(np.sum(direction ** 2))
    return _slice_along(direction, init_x, logprob, sigma, step_out, max_steps_out)
