"""CPU, world_size 2 over gloo: the sample sharding + single all-reduce reproduces the unsharded mean EI.
(The per-rank compute here is the oracle -- this test covers the host-side exchange logic only.)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import gp_oracle as O
    from spearmint_b200 import parallel
    from tests.helpers import hypers, load, sets
    g = load("opt_d8_m52")
    comp, pend, cand, vals = sets(g)
    hs = hypers(g)

    def local(idx):
        tot = np.zeros(cand.shape[0])
        for s in idx:
            tot += O.compute_ei("Matern52", hs[s], comp, pend, cand, vals)
        return torch.from_numpy(tot)

    mean_ei = parallel.sharded_mean_ei(local, len(hs)).numpy()
    assert parallel.shard(len(hs), rank, world) == list(range(rank, len(hs), world))
    if rank == 0:
        np.save(out, mean_ei)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_mean_ei_matches_unsharded(tmp_path):
    from tests.helpers import load
    out = str(tmp_path / "mean_ei.npy")
    mp.spawn(_worker, args=(2, 29531, out), nprocs=2, join=True)
    g = load("opt_d8_m52")
    np.testing.assert_allclose(np.load(out), g["overall_ei"].mean(axis=1), rtol=1e-9, atol=1e-15)


def test_shard_covers_all_samples_once():
    from spearmint_b200 import parallel
    for S in (1, 5, 40):
        for W in (1, 2, 4, 8):
            got = sorted(sum((parallel.shard(S, r, W) for r in range(W)), []))
            assert got == list(range(S))


def _worker_err(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spearmint_b200 import parallel
    res = []
    parallel.agree_on_error(None)                                  # nobody failed: returns on every rank
    res.append("ok")
    try:                                                           # only rank 1's shard is not positive definite
        parallel.agree_on_error(np.linalg.LinAlgError("rank 1: not positive definite") if rank == 1 else None)
        res.append("no-raise")
    except np.linalg.LinAlgError as e:
        res.append("raised:" + str(e))
    t = torch.ones(3)                                              # and the next collective still lines up
    parallel.allreduce_sum_(t)
    res.append(float(t[0]))
    with open(out + ".%d" % rank, "w") as fh:
        fh.write(repr(res))
    dist.barrier()
    dist.destroy_process_group()


def test_an_error_on_one_rank_is_raised_on_every_rank(tmp_path):
    """a LinAlgError in one rank's shard of hyper-samples must not leave the other ranks waiting in the EI
    all-reduce -- every rank raises (parallel.agree_on_error), and the collectives stay aligned afterwards."""
    out = str(tmp_path / "err")
    mp.spawn(_worker_err, args=(2, 29533, out), nprocs=2, join=True)
    r0, r1 = eval(open(out + ".0").read()), eval(open(out + ".1").read())
    assert r0[0] == "ok" and r1[0] == "ok"
    assert r0[1].startswith("raised:") and "another rank" in r0[1]
    assert r1[1] == "raised:rank 1: not positive definite"
    assert r0[2] == 2.0 and r1[2] == 2.0
