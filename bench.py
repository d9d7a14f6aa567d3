#!/usr/bin/env python
"""bench.py -- EI candidates/sec of the GP-EI chooser hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" is one pass of ``ei_over_hypers`` (chooser/GPEIOptChooser.py:331-341 of the reference) over the
whole candidate grid for all S hyper-samples: K build -> Cholesky -> explicit inverse + alpha -> cross-covariance
generator -> tcgen05 predict GEMM (3xFP16 scaled split) -> EI sweep (-> all-reduce of the per-candidate EI sum when
N > 1) -> argmax.

  value : M / step-time with X, candidates, values and hyper-samples already resident in HBM.
  e2e   : the same metric through the host-facing call (numpy in, (M,S) EI matrix out), host<->device copies
          inside the timed region.
  roofline     : the dominant kernel (tc::predict_tc_kernel, the N^2*M triangular-solve term as a GEMM against the
                 explicit inverse), algorithmic flops / its CUDA-event time vs the measured dense bf16 tensor peak
                 (MEASURED_PEAKS.json); traffic = DRAM bytes per launch from the ncu capture in profiles/.
  cpu_baseline : the oracle port (numpy/scipy, the reference's own operation sequence) timed on this box's
                 host cores on a bounded sample of the same workload and extrapolated linearly (the reference
                 loop is exactly linear in S and in M for fixed N).
  --impl reference : that CPU implementation as its own arm (rank 0 only).

Multi-GPU: hyper-samples are sharded round-robin over ranks (total work fixed -> "scaling": "strong"); the
only exchange is one NCCL all-reduce of M floats.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (D, N, M, S)           BASELINE.json configs / metric
    "headline": (32, 4096, 100000, 40),   # "EI candidates/sec (N=4096 obs, 100k cands, 40 hypers)"
    "c2": (8, 512, 10000, 10),
    "c3": (20, 2048, 50000, 20),
    "c5": (32, 8192, 100000, 40),
    "tiny": (4, 96, 2000, 4),
}
KIND = "Matern52"


def synth(D, N, M, S):
    """SURVEY.md 8(d) synthetic problem; the grid is RandomState(0).rand (not Sobol: generation time only)."""
    grid = np.random.RandomState(0).rand(N + M, D)
    perm = np.random.RandomState(0).permutation(N + M)
    complete, candidates = np.sort(perm[:N]), np.sort(perm[N:])
    comp, cand = grid[complete], grid[candidates]
    y = np.sin(3 * comp).sum(1) + 0.01 * np.random.RandomState(1).randn(N)
    vals = (y - y.mean()) / y.std()
    rs = np.random.RandomState(2)
    hs = [(0.1 * rs.randn(), 1e-3, float(np.exp(0.25 * rs.randn())), rs.uniform(0.3, 2.0, D)) for _ in range(S)]
    return comp, cand, vals, hs


def flops_per_pair(N, D):
    """algorithmic flops per (candidate, hyper-sample) pair."""
    return float(N) * N + 2.0 * N * D + 4.0 * N + 25.0 * N + 40.0


# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE predict-GEMM launch over a full 32768-candidate chunk
# of the headline workload on one GPU, from the `ncu --set full` capture summarised in profiles/ (bench.py never runs
# under a profiler).  A step has 3 such launches + a 1696-candidate tail; the per-launch average is reported.
NCU_TRAFFIC = {"headline": {"bytes_per_full_chunk_launch": 37.82e9, "chunk_cands": 32768,
                            "source": "profiles/r01_predict_tc_final_ncu.md"}}


def ncu_traffic(workload, impl, world, M, launches_per_step):
    t = NCU_TRAFFIC.get(workload)
    if t is None or impl != "tc" or world != 1:
        return None
    return t["bytes_per_full_chunk_launch"] * (float(M) / t["chunk_cands"]) / max(1, launches_per_step)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor_sustained=d["bf16_tflops_sustained"],
                    src="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(object):
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, n in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_port_sample(D, N, M, S, budget_cands=None, threads=None):
    """Times the oracle port on host cores on a bounded sample and extrapolates to the full (M, S) pass.

    The reference loop is linear in S (OPT:333-340) and, for fixed N, in M (every stage after the Cholesky is
    per-candidate), so   t_full = S * (t_fixed + t_cand * M / M_cpu)."""
    from oracle import gp_oracle as O
    import scipy.linalg as spla
    comp, cand, vals, hs = synth(D, N, M, S)
    budget_cands_auto = budget_cands is None
    if budget_cands is None:   # first guess; grown below if the host is fast
        budget_cands = int(max(500, min(M, 2.0e11 / (float(N) * N + 60.0 * N * D))))
    Mc = min(M, budget_cands)
    h = hs[0]
    t0 = time.perf_counter()
    mean, noise, amp2, ls = h
    K = O.cov(KIND, amp2, ls, comp) + noise * np.eye(N)
    L = spla.cholesky(K, lower=True)
    alpha = spla.cho_solve((L, True), vals - mean)
    t_fixed = time.perf_counter() - t0
    def candidates_part(mc):
        t0 = time.perf_counter()
        Kx = O.cov(KIND, amp2, ls, comp, cand[:mc])
        beta = spla.solve_triangular(L, Kx, lower=True)
        m = np.dot(Kx.T, alpha) + mean
        v = amp2 * (1 + 1e-6) - np.sum(beta ** 2, axis=0)
        e = O._ei_from_moments(np.min(vals), m, np.sqrt(v))
        return time.perf_counter() - t0, e

    t_cand, ei = candidates_part(Mc)
    if budget_cands_auto and t_cand < 5.0 and Mc < M:      # many-core hosts: grow the sample to ~10 s of CPU work
        Mc = int(min(M, 50000, max(Mc, Mc * 10.0 / max(t_cand, 1e-3))))   # (<= 1.6 GB per N x Mc float64 temporary)
        t_cand, ei = candidates_part(Mc)
    t_full = S * (t_fixed + t_cand * (float(M) / Mc))
    return dict(value=M / t_full, t_full_s=t_full, t_fixed_s=t_fixed, t_cand_s=t_cand, Mc=Mc,
                sample="1 hyper-sample, N=%d, %d of %d candidates; extrapolated linearly to S=%d, M=%d "
                       "(t_fixed=%.2fs K+chol+alpha, t_cand=%.2fs)" % (N, Mc, M, S, M, t_fixed, t_cand),
                checksum=float(np.sum(ei)))


def run_reference_arm(args, D, N, M, S):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    vals = []
    for it in range(args.warmup + args.steps):
        r = cpu_port_sample(D, N, M, S, budget_cands=args.cpu_cands)
        if it >= args.warmup:
            vals.append(r)
    t = float(np.mean([r["t_full_s"] for r in vals]))
    value = M / t
    out = {"impl": "reference", "metric": "EI candidates/sec", "value": value, "unit": "candidates/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "%s D=%d N=%d M=%d S=%d %s" % (args.workload, D, N, M, S, KIND),
                      "note": "oracle port of the reference numpy/scipy path (the py2 reference cannot travel); "
                              "each step times a bounded sample and extrapolates linearly"},
           "cpu_baseline": {"value": value, "unit": "candidates/s", "cores": cores, "kind": "port",
                            "sample": vals[-1]["sample"]},
           "e2e": {"value": value, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------- GPU arm
def run_b200_arm(args, D, N, M, S):
    import torch
    import torch.distributed as dist
    from spearmint_b200 import _lib, parallel
    from spearmint_b200.engine import GPEIEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: the product arm needs a CUDA device (B200, sm_100a) -- there is no CPU fallback; "
                         "the CPU arm is `--impl reference`")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
    eng = GPEIEngine(device="cuda:%d" % local, dtype=torch.float32)
    comp, cand, vals, hs = synth(D, N, M, S)
    mine = parallel.shard(S, rank, world)
    hs_local = [hs[s] for s in mine]
    Sl = len(hs_local)

    # ---- resident inputs for the `value` leg
    res = dict(X=eng.to_dev(comp), C=eng.to_dev(cand), y=eng.to_dev(vals), best=float(vals.min()),
               hb=eng.hypers(hs_local, KIND) if Sl else None)
    ldm = ((M + 127) // 128) * 128

    def step_resident():
        if Sl:
            _, ei_sum, _ = eng.ei_over_hypers_device(KIND, hs_local, None, None, None, None, want_matrix=False,
                                                     inputs_on_device=res)
        else:
            ei_sum = torch.zeros((ldm,), dtype=torch.float64, device=eng.device)
        parallel.allreduce_sum_(ei_sum)
        idx, _ = eng.topk(ei_sum, M, 1)
        return idx

    def step_e2e():
        # host numpy in -> H2D -> path -> D2H of this rank's (M, S_local) EI columns + all-reduced argmax
        if Sl:
            ei, ei_sum, _ = eng.ei_over_hypers_device(KIND, hs_local, comp, None, cand, vals, want_matrix=True)
            host = ei[:, :M].t().contiguous().cpu()
        else:
            ei_sum = torch.zeros((ldm,), dtype=torch.float64, device=eng.device)
            host = None
        parallel.allreduce_sum_(ei_sum)
        idx, _ = eng.topk(ei_sum, M, 1)
        return host, int(idx.cpu()[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    L = _lib.lib()

    def timed(fn, steps, with_timers=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.timers = {} if with_timers else None
        L.smk_timing_enable(1 if with_timers else 0)
        l0 = _lib.lib().smk_launch_count()
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - w0
        ms = e0.elapsed_time(e1)
        launches = _lib.lib().smk_launch_count() - l0
        stages = eng.stage_ms()
        eng.timers = None
        if not with_timers:
            pass
        t = torch.tensor([ms, wall * 1e3, float(launches)], dtype=torch.float64, device=eng.device)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            ms, wall_ms, launches = float(tmax[0]), float(tmax[1]), int(tsum[2])
        else:
            wall_ms = wall * 1e3
        return ms, wall_ms, launches, stages, out

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, _, launches, stages, idx = timed(step_resident, args.steps, with_timers=True)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = M / (ms_per_step * 1e-3)

    # kernel spans of the resident leg are read below; freeze them before the e2e leg runs
    span = {}
    for nm in ("predict_tc_kernel", "predict_kernel", "kxt_kernel", "trtri_kernel", "linv_pack_f16"):
        c2 = ctypes.c_int(0)
        span[nm] = (L.smk_timing_ms(nm.encode(), ctypes.byref(c2)), c2.value)

    # ---- e2e leg (host buffers through the public call), device-event timed around host work too
    step_e2e()
    _, wall_ms, _, _, (host, best_idx) = timed(step_e2e, max(1, min(args.steps, 3)))
    e2e_ms = wall_ms / max(1, min(args.steps, 3))
    esz = 4
    h2d = esz * (comp.size + cand.size + vals.size) + esz * Sl * (D + 3)
    d2h = 8 * M * Sl + 4       # the (M, S_local) EI matrix is float64 (tail values), + the argmax index

    if rank == 0:
        pk = peaks()
        pairs_local = float(M) * Sl
        impl = eng.predict_impl
        kname = "predict_tc_kernel" if impl == "tc" else "predict_kernel"
        kms_total, cnt_v = span[kname]
        kms_step = kms_total / args.steps                       # all launches of the dominant kernel in one step
        n_launch_step = max(1, cnt_v // args.steps)
        # algorithmic flops handled by that kernel per step: the N^2 triangular-solve term for the tcgen05
        # kernel (the cross-covariance is a separate kernel there), the whole per-pair count for the fused SIMT kernel
        alg = (float(N) * N if impl == "tc" else flops_per_pair(N, D)) * pairs_local
        achieved = alg / (kms_step * 1e-3) / 1e12 if kms_step > 0 else None
        other = {}
        for nm in ("kxt_kernel", "trtri_kernel", "linv_pack_f16"):
            if span[nm][1]:
                other[nm + "_ms_per_step"] = span[nm][0] / args.steps
        roof = {"kernel": "smk::tc::predict_tc_kernel (tcgen05.mma kind::f16, 3xFP16 scaled split, fp32 accumulate)" if impl == "tc"
                else "smk::predict_kernel<float> (SIMT FMA)",
                "bound": "tensor", "achieved": achieved, "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                "frac": (achieved / pk["tensor_sustained"]) if achieved else None,
                "traffic": ncu_traffic(args.workload, impl, world, M, n_launch_step) if Sl == 40 else None,
                "peak_source": pk["src"] + " bf16 dense, sustained (kernel timed inside a long step)",
                "algorithmic_flops_per_launch": alg / n_launch_step, "launches_per_step": n_launch_step,
                "kernel_ms_per_launch": kms_step / n_launch_step, "kernel_ms_per_step": kms_step,
                "note": "fp32-accurate results need 3 fp16 MMAs per product (hi*hi + hi*lo + lo*hi, operands scaled by exact "
                        "powers of two): 3 bf16-equivalent tensor flops per algorithmic flop, so 1/3 = 0.333 of the "
                        "bf16 peak is this formulation's ceiling"
                        if impl == "tc" else "float32 SIMT FMA path (B200 fp32 vector peak ~74 TFLOP/s)",
                "stage_ms_per_step": dict({k: v / args.steps for k, v in stages.items()}, **other)}
        cpu = cpu_port_sample(D, N, M, S, budget_cands=args.cpu_cands) if world == 1 and not args.no_cpu else None
        out = {"metric": "EI candidates/sec", "value": value, "unit": "candidates/s", "n_gpus": world,
               "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic",
               "config": {"workload": "%s D=%d N=%d M=%d S=%d %s" % (args.workload, D, N, M, S, KIND),
                          "parallelism": "hyper-samples round-robin over %d GPU(s), one all-reduce of M floats" % world,
                          "l2": "working set (factors %.1f GB per rank) >> 126 MB L2, no flush needed"
                                % (Sl * (((N + 127) // 128) * 128) ** 2 * 4 / 1e9),
                          "pairs_per_sec": value * S, "argmax": int(idx.cpu()[0]), "argmax_e2e": best_idx},
               "e2e": {"value": M / (e2e_ms * 1e-3), "unit": "candidates/s", "ms_per_step": e2e_ms,
                       "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "api": "GPEIEngine.ei_over_hypers_device(host numpy) -> (M,S_local) EI matrix on host"},
               "gpu_launches": launches, "clocks": clocks, "roofline": roof}
        if cpu is not None:
            out["cpu_baseline"] = {"value": cpu["value"], "unit": "candidates/s", "cores": os.cpu_count(),
                                   "kind": "port", "sample": cpu["sample"]}
        if world == 1 and not args.no_next:
            out["next"] = next_wall_ms(args)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def next_wall_ms(args):
    """Second half of BASELINE.json's metric: chooser.next() wall-ms through the plugin API (MCMC chain with the GPU
    float64 log-likelihood -> grid pass -> L-BFGS refinement with cached factors -> grid pass).  Measured at the `c2`
    configuration (D=8, N=512, 10k candidates, 10 hyper-samples, reference defaults burnin=100, grid_subset=20); the CPU
    figure (same host logic on the oracle numerics = a port of the reference's next()) only with --next-cpu: it takes
    ~20 s per call."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import next_bench
    D, N, M, S = WORKLOADS["c2"]
    calls = next_bench.run("gpu", D, N, M, S, burnin=100, calls=3, grid_subset=20)
    out = {"workload": "c2 D=%d N=%d M=%d S=%d burnin=100 grid_subset=20" % (D, N, M, S),
           "ms_first_call_with_burnin": calls[0]["ms"], "ms_steady": float(np.mean([c["ms"] for c in calls[1:]])),
           "unit": "ms", "refine_evals": calls[-1]["refine_evals"]}
    if args.next_cpu:
        cpu = next_bench.run("cpu", D, N, M, S, burnin=100, calls=2, grid_subset=20)
        out["cpu_port_ms_steady"] = cpu[-1]["ms"]
        out["cpu_port_ms_first_call"] = cpu[0]["ms"]
        out["cpu_cores"] = os.cpu_count()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="headline", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-cands", type=int, default=None, help="candidates in the CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--samples", type=int, default=None, help="override S (experiments only)")
    ap.add_argument("--no-next", action="store_true", help="skip the chooser.next() wall-ms leg")
    ap.add_argument("--next-cpu", action="store_true", help="also time the CPU port of next() (slow)")
    args = ap.parse_args()
    D, N, M, S = WORKLOADS[args.workload]
    if args.samples:
        S = args.samples
    if args.impl == "reference":
        run_reference_arm(args, D, N, M, S)
    else:
        run_b200_arm(args, D, N, M, S)


if __name__ == "__main__":
    main()
