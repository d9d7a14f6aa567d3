"""Shared helpers for the parity tests (fixture loading; no compute)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

OPT_CASES = ["opt_branin2d", "opt_d8_m52", "opt_d8_m52_pend", "opt_d5_ardse", "opt_d4_m32_pend",
             "opt_d3_se", "opt_d1_m52"]
PSEC_CASES = ["psec_d4", "psec_d3_pend"]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def hypers(g, prefix="hs"):
    """list of (mean, noise, amp2, ls) tuples, the reference's hyper_samples layout (OPT:628)."""
    return [(float(g[prefix + "_mean"][s]), float(g[prefix + "_noise"][s]), float(g[prefix + "_amp2"][s]),
             np.array(g[prefix + "_ls"][s], dtype=float)) for s in range(len(g[prefix + "_mean"]))]


def sets(g):
    grid, values = g["grid"], g["values"]
    comp = grid[g["complete"]]
    cand = grid[g["candidates"]]
    pend = grid[g["pending"]] if g["pending"].size else np.zeros((0, grid.shape[1]))
    return comp, pend, cand, values[g["complete"]]
