"""CPU, world_size 2 over gloo: the N>1 host path -- trial slab bounds, the det-term split and the single
all-reduce of the partial (loglik, gradient) vector reproduce the unsharded oracle result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_bounds_cover_all_trials():
    from gpcsd_b200.parallel import shard_bounds
    for n in (0, 1, 7, 50, 2000, 2003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _partial_eval(om, lfp_slab, ntot, frac):
    """What one rank contributes: data terms from its slab, trial-independent terms weighted by `frac`
    (mirrors KronEngine.loglik_grad's host assembly, using the oracle as the per-slab evaluator)."""
    from oracle import gpcsd_oracle as O
    n = lfp_slab.shape[2]
    ll_n, g_n = O.loglik_and_grad(om, lfp_slab)              # = n*det + data(slab)
    ll_0, g_0 = O.loglik_and_grad(om, lfp_slab[:, :, :0])    # = 0 (no trials): det scaling check
    assert ll_0 == 0.0
    # det part per trial from a one-trial zero LFP: loglik(zeros, 1 trial) = -0.5 sum log D
    zero = np.zeros(lfp_slab.shape[:2] + (1,))
    ll_det, g_det = O.loglik_and_grad(om, zero)
    data_ll, data_g = ll_n - n * ll_det, g_n - n * g_det
    return np.concatenate([[data_ll + frac * ntot * ll_det], data_g + frac * ntot * g_det])


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gpcsd_b200.parallel import TrialShard
    from oracle import gpcsd_oracle as O, synth
    x, t = synth.geometry_1d(24, 30)
    om = synth.model_1d(x, t, sig2n=1e-2)
    lfp = synth.matched_lfp(om, 11, 5)
    sh = TrialShard(True)
    assert (sh.rank, sh.world) == (rank, world)
    lo, hi = sh.bounds(lfp.shape[2])
    part = _partial_eval(om, lfp[:, :, lo:hi], lfp.shape[2], sh.det_fraction())
    tot = sh.allreduce_sum(part)
    ntot = sh.allreduce_sum(np.array([float(hi - lo)]))
    ll, g = O.loglik_and_grad(om, lfp)
    ok = abs(tot[0] - ll) <= 1e-10 * abs(ll) and np.max(np.abs(tot[1:] - g) / np.abs(g)) < 1e-9 and int(ntot[0]) == 11
    q.put((rank, bool(ok), float(tot[0]), float(ll)))
    dist.destroy_process_group()


def test_world2_gloo_allreduce_matches_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res


def test_trialshard_disabled_by_default():
    from gpcsd_b200.parallel import TrialShard
    sh = TrialShard(None)
    assert not sh.enabled and sh.world == 1 and sh.bounds(10) == (0, 10)
    v = np.array([1.0, 2.0])
    assert np.array_equal(sh.allreduce_sum(v), v)
    with pytest.raises(RuntimeError):
        TrialShard(True)
