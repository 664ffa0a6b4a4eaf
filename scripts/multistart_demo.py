#!/usr/bin/env python
"""BASELINE.json configs[3]: GPCSD1D hyperparameter fit with 64 multi-start restarts; restarts sharded over the ranks
(`distributed_restarts=True`: every rank holds the full LFP, no collective until the final gather of 64 results).

    python scripts/multistart_demo.py [nt] [ntrials] [n_restarts]
    torchrun --nproc-per-node 8 scripts/multistart_demo.py 50 50 64
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    args = [a for a in sys.argv[1:] if a != "--sequential"]
    nt = int(args[0]) if len(args) > 0 else 50
    ntrials = int(args[1]) if len(args) > 1 else 50
    n_restarts = int(args[2]) if len(args) > 2 else 64
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=600))
    import __graft_entry__ as ge
    ge.ensure_built()
    from gpcsd_b200.engine import KronEngine
    from gpcsd_b200.gpcsd1d import GPCSD1D
    from oracle import synth   # input generator only
    x, t = synth.geometry_1d(24, nt)
    om = synth.model_1d(x, t, sig2n=1e-2)
    lfp = synth.matched_lfp(om, ntrials, 1000)
    count = {"n": 0, "calls": 0}
    orig = KronEngine.loglik_grad_batch

    def counted(self, hps, want_grad=True):
        hps = list(hps)
        count["n"] += len(hps)                     # evaluations (restart x step)
        count["calls"] += 1                        # native calls (one per lock step)
        return orig(self, hps, want_grad)
    KronEngine.loglik_grad_batch = counted
    orig_t = KronEngine.loglik_grad_thetas

    def counted_t(self, thetas, template, want_grad=True):
        count["n"] += len(thetas)
        count["calls"] += 1
        return orig_t(self, thetas, template, want_grad)
    KronEngine.loglik_grad_thetas = counted_t
    lockstep = "--sequential" not in sys.argv
    sys.argv = [a for a in sys.argv if a != "--sequential"]
    np.random.seed(1)
    m = GPCSD1D(lfp, x, t, distributed_restarts=(world > 1))
    m.loglik()                                             # upload + warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.fit(n_restarts=n_restarts, lockstep=lockstep)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = torch.tensor([float(count["n"]), dt], dtype=torch.float64, device="cuda")
    if world > 1:
        tot = n.clone()
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        mx = n.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        nev, wall = float(tot[0]), float(mx[1])
    else:
        nev, wall = float(n[0]), dt
    if rank == 0:
        p = m.extract_model_params()
        print("multi-start fit 24x%dx%d (%s): %d restarts on %d GPU(s): %d loglik+grad evaluations in %.2f s -> %.0f evals/s; "
              "R %.1f ell %.1f sig2n %.4f" % (nt, ntrials, "lock step" if lockstep else "scipy per restart", n_restarts, world, nev, wall,
                                               nev / wall, p['R'], p['spatial_ell'], p['sig2n']))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
