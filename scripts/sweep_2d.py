#!/usr/bin/env python
"""BASELINE.json configs[4]: GPCSD2D scaling sweep -- loglik+grad time vs (time points, trials) on the
Neuropixels geometry (384 channels, 30 x 120 quadrature), reported as evals/s and as the algorithmic
FLOP rate F_lg = 4 N nx nt (nx + nt) against the in-run cuBLAS DGEMM rate.

    python scripts/sweep_2d.py [--nt 100 250 500 1000 2000] [--trials 1000 4000] [--out profiles/r01_sweep_2d.json]
Under torchrun the trials are sharded over the ranks (strong scaling of one evaluation).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def neuropixels_geometry(nch, nt, dt=0.4):
    ch = np.arange(nch)
    xs = np.array([16.0, 48.0, 0.0, 32.0])[ch % 4]
    ys = 20.0 * np.floor(ch / 2)
    return np.stack([xs, ys], axis=1), (dt * np.arange(nt, dtype=np.float64))[:, None]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nt", type=int, nargs="+", default=[100, 250, 500, 1000, 2000])
    ap.add_argument("--trials", type=int, nargs="+", default=[1000, 4000])
    ap.add_argument("--nch", type=int, default=384)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    import __graft_entry__ as ge
    ge.ensure_built()
    import scipy.special
    from gpcsd_b200.engine import HyperParams, KronEngine

    def gl(a, b, n):
        u, w = scipy.special.roots_legendre(n)
        return 0.5 * (u + 1) * (b - a) + a, 0.5 * (b - a) * w

    # cuBLAS DGEMM reference rate
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    best = 1e9
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        if i >= 2:
            best = min(best, e0.elapsed_time(e1))
    peak = 2.0 * n ** 3 / (best * 1e-3) * 1e-12
    del a, b
    rows = []
    for nt in args.nt:
        X, t = neuropixels_geometry(args.nch, nt)
        g1, w1 = gl(-16.0, 64.0, 30)
        g2, w2 = gl(-100.0, float(X[:, 1].max()) + 100.0, 120)   # fit_gpcsd2d.py:86-90
        eng = KronEngine(2, X, t, dict(gl_x1=g1, gl_w1=w1, gl_x2=g2, gl_w2=w2), group=(True if world > 1 else None))
        hp = HyperParams(R=100.0, ells=(40.0, 200.0), temporal=[(0, 5.0, 1e-10), (1, 1.0, 1.4e-10)], sig2n=0.5, eps=1.0)
        for N in args.trials:
            nloc = N // world
            gen = torch.Generator(device=dev)
            gen.manual_seed(1234 + rank)
            Y = torch.randn((args.nch, nt, nloc), dtype=torch.float64, device=dev, generator=gen)
            eng.set_lfp(Y, local=(world > 1))
            del Y
            eng.loglik_grad(hp)
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.reps):
                if world > 1:
                    dist.barrier()
                t0 = time.perf_counter()
                ll, g = eng.loglik_grad(hp)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            dt = float(np.median(ts))
            if world > 1:
                tt = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            flg = 4.0 * N * args.nch * nt * (args.nch + nt)
            rows.append({"nch": args.nch, "nt": nt, "trials": N, "n_gpus": world, "ms_per_eval": 1e3 * dt, "evals_per_s": 1.0 / dt,
                         "algorithmic_tflops": flg / dt * 1e-12, "frac_of_dgemm_per_gpu": flg / dt * 1e-12 / (peak * world),
                         "lfp_gb_per_gpu": args.nch * nt * nloc * 8e-9, "loglik": float(ll)})
            if rank == 0:
                print(json.dumps(rows[-1]), flush=True)
            eng.Y = None
            eng._ws = {k: v for k, v in eng._ws.items() if k[0] not in ("Z", "Bm")}
            for pl in eng._plans.values():            # the native plan's workspace is sized by the trial count: release it
                pl.ws, pl._bound = None, None
            torch.cuda.empty_cache()
        del eng
        torch.cuda.empty_cache()
    if rank == 0 and args.out:
        json.dump({"dgemm_tflops": peak, "rows": rows}, open(args.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
