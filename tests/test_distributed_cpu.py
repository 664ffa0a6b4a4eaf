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
    # the device-vector entry point falls back to the host path on gloo, and joins a CollectiveOrder rotation transparently
    from gpcsd_b200.parallel import CollectiveOrder
    order = CollectiveOrder(1)
    sh.set_order(order, 0)
    order.start()
    tot_dev = sh.allreduce_device(torch.from_numpy(part.copy()))
    order.stop()
    assert np.array_equal(tot_dev, tot)
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


class _StubModel:
    """GPCSDModelBase._fit driven by a cheap analytic objective (no CUDA): checks the restart-sharding logic."""

    def __new__(cls, restart_group):
        from gpcsd_b200._model import GPCSDModelBase
        from gpcsd_b200.priors import GPCSDHalfNormalPrior, GPCSDInvGammaPrior

        class M(GPCSDModelBase):
            DIM = 1
            SPATIAL_ELL_KEYS = ('ell',)

            def _get_engine(self):
                return None

            def _pure_objective(self, engine, fix_R):
                def fun(tparams):
                    tp = np.asarray(tparams)
                    target = np.linspace(-0.5, 0.5, tp.size)
                    # two basins so that different starts end in different local minima
                    f = np.sum((tp - target) ** 2 * ((tp - target - 1.5) ** 2 + 0.3))
                    g = 2 * (tp - target) * ((tp - target - 1.5) ** 2 + 0.3) + (tp - target) ** 2 * 2 * (tp - target - 1.5)
                    return float(f), g
                return fun

            def _batched_objective(self, engine, fix_R):
                one = self._pure_objective(engine, fix_R)

                def fun(X, idx):
                    vals = [one(x) for x in np.atleast_2d(X)]
                    return np.array([v[0] for v in vals]), np.array([v[1] for v in vals])
                return fun

        m = M()
        ig = GPCSDInvGammaPrior(); ig.set_params(50.0, 500.0)

        class SC:
            params = {'ell': {'value': 200.0, 'prior': ig, 'min': 10.0, 'max': 5000.0}}

        class TC:
            KIND = 0
            def __init__(self):
                tg = GPCSDInvGammaPrior(); tg.set_params(1.0, 40.0)
                self.params = {'ell': {'value': 5.0, 'prior': tg, 'min': 0.1, 'max': 500.0},
                               'sigma2': {'value': 1.0, 'prior': GPCSDHalfNormalPrior(1.0), 'min': 1e-8, 'max': np.inf}}
        m.spatial_cov = SC()
        m.temporal_cov_list = [TC(), TC()]
        m.R = {'value': 100.0, 'prior': ig, 'min': 10.0, 'max': 5000.0}
        m.sig2n = {'value': 0.1, 'prior': GPCSDHalfNormalPrior(0.1), 'min': 1e-8, 'max': 0.5}
        m._restart_group = restart_group
        return m


def _fit_stub(restart_group, seed=11, n_restarts=6, lockstep=False):
    np.random.seed(seed)
    m = _StubModel(restart_group)
    m._fit(n_restarts, 'L-BFGS-B', False, False, {'maxiter': 200, 'gtol': 1e-10}, n_workers=1, lockstep=lockstep)
    return np.array([m.R['value'], m.spatial_cov.params['ell']['value'], m.sig2n['value']] +
                    [tc.params[k]['value'] for tc in m.temporal_cov_list for k in ('ell', 'sigma2')])


def _restart_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gpcsd_b200.parallel import RestartShard
    sh = RestartShard(True)
    assert [i for i in range(6) if sh.mine(i)] == list(range(rank, 6, world))
    q.put((rank, np.concatenate([_fit_stub(True), _fit_stub(True, lockstep=True)])))
    dist.destroy_process_group()


def test_restart_sharding_world2_matches_unsharded():
    """fit() with restarts sharded over 2 ranks ends at exactly the parameters of the unsharded fit."""
    single = np.concatenate([_fit_stub(None), _fit_stub(None, lockstep=True)])      # per-restart scipy path, lock-step path
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_restart_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert np.array_equal(res[0], res[1])
    assert np.allclose(res[0], single, rtol=0, atol=0)
