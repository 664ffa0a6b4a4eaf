#!/usr/bin/env python
"""Row (f) of SURVEY.md section 8: the fit() driver on configs[1]-shaped data -- one probe, 24 ch x 500 t x 2000 trials,
per-electrode noise (30 hyperparameters), set up like auditory_lfp/fit_gpcsd_baseline.py:80-96.  Prints wall time,
objective evaluations and the recovered hyperparameters."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if a != "--sequential"]
    n_restarts = int(args[0]) if len(args) > 0 else 3
    n_workers = int(args[1]) if len(args) > 1 else 2
    torch.cuda.set_device(0)
    import __graft_entry__ as ge
    ge.ensure_built()
    m = bench.make_models(torch.device("cuda", 0), 1, torch)[0]
    truth = m.extract_model_params()
    from gpcsd_b200.engine import KronEngine
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
    np.random.seed(0)
    f_true = m.obj_fun(np.log(np.array([truth['R'] / 100, truth['spatial_ell'] / 100] +
                                       [v for pair in zip(truth['temporal_ell_list'], truth['temporal_sigma2_list']) for v in pair] +
                                       list(truth['sig2n']))))
    t0 = time.perf_counter()
    m.fit(n_restarts=n_restarts, verbose=True, n_workers=n_workers, lockstep=lockstep)
    dt = time.perf_counter() - t0
    fit = m.extract_model_params()
    tp = np.log(np.array([fit['R'] / 100, fit['spatial_ell'] / 100] +
                         [v for pair in zip(fit['temporal_ell_list'], fit['temporal_sigma2_list']) for v in pair] + list(fit['sig2n'])))
    f_fit = m.obj_fun(tp)
    print("fit (%s): %d restarts, %d objective+gradient evaluations in %d native calls, %.2f s wall (%.2f ms per evaluation)"
          % ("lock step" if lockstep else "scipy per restart, %d workers" % n_workers, n_restarts, count["n"], count["calls"], dt,
             1e3 * dt / max(count["n"], 1)))
    print("nll at generating parameters %.3f, at fitted parameters %.3f" % (f_true, f_fit))
    print("R %.1f -> %.1f | ell %.1f -> %.1f | ell_t %s -> %s" % (truth['R'], fit['R'], truth['spatial_ell'], fit['spatial_ell'],
                                                              np.round(truth['temporal_ell_list'], 2), np.round(fit['temporal_ell_list'], 2)))
    print("median sig2n %.4f -> %.4f" % (np.median(truth['sig2n']), np.median(fit['sig2n'])))


if __name__ == "__main__":
    main()
