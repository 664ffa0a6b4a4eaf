#!/usr/bin/env python
"""bench.py -- GPCSD loglik+grad throughput on B200 (BASELINE.json metric, configs[1]).

Workload (configs[1]): GPCSD1D auditory-shaped synthetic LFP -- 2 probes x 24 channels, 500 time points
(1 ms grid), 2000 trials per probe per GPU, integration bounds a=-200, b=2600, ngl=100, temporal list
[SE, Matern-1/2], per-electrode noise (P = 30 hyperparameters), set up like
auditory_lfp/fit_gpcsd_baseline.py:80-89 of the reference.  A STEP is one marginal log-likelihood +
hyperparameter-gradient evaluation for EACH of the two probes (2 evals); hyperparameters change every
step (theta_true + 0.1 N(0,1), the L-BFGS access pattern).  The unit "eval" is one loglik+grad over one
24 x 500 x 2000 trial block.

  value : device-resident -- each probe's LFP block already sits in HBM (how fit() runs: data uploaded
          once, thousands of evaluations).  Everything else is inside the timed region: covariance build,
          both eigendecompositions, projections, gradient SYRKs, device->host read of (ll, grad), host
          assembly, and the all-reduce at N > 1.
  e2e   : the same evaluations through the public API (GPCSD1D.update_lfp + GPCSD1D.obj_fun_and_grad) with
          the LFP block copied from PINNED HOST memory every step and the result read back.
Multi-GPU (torchrun, one rank per GPU): weak scaling -- every rank holds its own 2000-trial slab per probe,
the model sees 2000*N trials, one all-reduce of P+1 doubles per evaluation; value = N * 2 * K / time.

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, numpy + BLAS threads)
on a bounded sample of the same workload on the box's host cores (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv[1:] or "--impl=reference" in sys.argv[1:]:
    # The CPU arm uses every host core whatever launched it: torch.distributed.run exports OMP_NUM_THREADS=1 to its
    # workers, which halved the round-1 denominator at N > 1.  Must happen before numpy loads its BLAS.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

# Throughput setting of the library for hosts that evaluate several models concurrently (the two probes of the concurrent
# arm): the GEMM phase of an evaluation leaves 16 SMs to the other model's eigensolver clusters (costs the
# bit-for-bit reproducibility across concurrency patterns that the default keeps -- results agree to rounding).  Only the
# overlapped (two-phase) calls are affected; the serial pass and every single-model line run the default path.
os.environ.setdefault("GPCSD_GEMM_RESERVE", "16")

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX, NT, NTRIALS, NGL = 24, 500, 2000, 100
A_LO, B_HI = -200.0, 2600.0
NPROBES = 2
METRIC = "gpcsd_loglik_grad_evals_per_s"
UNIT = "evals/s"


def true_hyper(probe):
    """theta_true of SURVEY.md 8d (sim_from_gp_1D.py:41-47 family), per-electrode noise 1e-2*exp(.3 N(0,1))."""
    rng = np.random.default_rng(2000 + probe)
    return dict(R=100.0, ell=200.0, se=(20.0, 0.5), matern=(5.0, 0.7), sig2n=1e-2 * np.exp(0.3 * rng.standard_normal(NX)))


def geometry():
    x = np.linspace(0.0, 2300.0, NX)[:, None]
    t = np.arange(NT, dtype=np.float64)[:, None]
    return x, t


# ----------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ----------------------------------------------------------------------------------------------------
def blas_threads():
    """(threads the BLAS behind numpy will use, library name) -- printed next to every CPU number."""
    try:
        from threadpoolctl import threadpool_info
        infos = [i for i in threadpool_info() if i.get("user_api") == "blas"]
        if infos:
            return int(max(i.get("num_threads", 1) for i in infos)), str(infos[0].get("internal_api", "blas"))
    except Exception:
        pass
    return int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1)), "unknown"


_CPU_CASE = {}


def cpu_case(ntrials=NTRIALS):
    """One probe block of the workload for the CPU arm: the oracle model of probe 0 and a full 24 x 500 x ntrials block
    (values do not change the arithmetic performed)."""
    if ntrials not in _CPU_CASE:
        from oracle import synth
        x, t = geometry()
        th = true_hyper(0)
        om = synth.model_1d(x, t, a=A_LO, b=B_HI, ngl=NGL, sig2n=th["sig2n"])
        lfp = np.random.default_rng(7).standard_normal((NX, NT, ntrials))
        _CPU_CASE[ntrials] = (om, lfp)
    return _CPU_CASE[ntrials]


def cpu_eval_seconds(ntrials=NTRIALS, reps=1, seed0=100):
    """Time the CPU restatement (oracle.gpcsd_oracle.loglik_and_grad: the reference's covariance build, two LAPACK eigh,
    Kronecker projection of every trial and the closed-form gradient) on ONE FULL probe block -- no extrapolation.
    Returns the list of seconds per evaluation."""
    from oracle import gpcsd_oracle as O
    from oracle import synth
    om, lfp = cpu_case(ntrials)
    ts = []
    for r in range(reps):
        om_r = synth.perturbed(om, seed0 + r)           # new hyperparameters every evaluation, like the GPU arm
        t0 = time.perf_counter()
        O.loglik_and_grad(om_r, lfp)
        ts.append(time.perf_counter() - t0)
    return ts


def cpu_predict_baseline(sample_trials=250):
    """Kronecker-form numpy port of predict (oracle.predict_kron) on a bounded sample; the reference's own dense
    (nx nt)^2 formulation needs 36 s and 4.8 GB at this shape and is not timed here."""
    from oracle import gpcsd_oracle as O
    from oracle import synth
    x, t = geometry()
    om = synth.model_1d(x, t, a=A_LO, b=B_HI, ngl=NGL, sig2n=true_hyper(0)["sig2n"])
    lfp = np.random.default_rng(8).standard_normal((NX, NT, sample_trials))
    O.predict_kron(om, lfp[:, :, :16], x, t, "csd")
    t0 = time.perf_counter()
    O.predict_kron(om, lfp, x, t, "csd")
    dt = time.perf_counter() - t0
    nthr, blas = blas_threads()
    return {"value": sample_trials / dt, "unit": "trials/s", "cores": nthr, "host_cpus": os.cpu_count(), "kind": "port",
            "sample": "oracle predict_kron (Kronecker-form numpy port) on 24x500x%d trials (linear in trials; not "
                      "extrapolated: the value is the sample's own rate)" % sample_trials}


def run_reference(args, rank):
    """CPU arm: the reference algorithm's restatement (oracle/, numpy + all BLAS threads) on the SAME unit of work as the
    GPU arm -- one step = one loglik+grad over a full 24 x 500 x 2000 block for each of the two probes, new hyperparameters
    every evaluation.  Nothing is extrapolated: every timed evaluation runs in full.  The unit ("eval" = one 2000-trial
    block) does not depend on the GPU count, so the same denominator applies at every N (weak scaling multiplies the blocks,
    not the block size).  Rank 0 only."""
    if rank != 0:
        return
    nthr, blas = blas_threads()
    cpu_case()
    for w in range(args.warmup):
        cpu_eval_seconds(reps=NPROBES, seed0=50 + NPROBES * w)
    t0 = time.perf_counter()
    n = 0
    for s in range(args.steps):
        n += len(cpu_eval_seconds(reps=NPROBES, seed0=100 + NPROBES * s))
    total = time.perf_counter() - t0
    ms = 1e3 * total / args.steps
    value = n / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthr, "host_cpus": os.cpu_count(), "blas": blas,
                             "kind": "port",
                             "sample": "oracle loglik_and_grad (numpy/LAPACK restatement of loglik + closed-form gradient; "
                                       "the reference is pure Python and HIPS autograd is not installable, so kind = port) "
                                       "on FULL 24x500x2000 blocks, %d evaluations timed, none extrapolated" % n},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "BASELINE.json configs[1]: GPCSD1D auditory-shaped, 2 probes x 24 ch x 500 t x 2000 trials per GPU, "
                        "per-electrode noise (P=30), a=-200 b=2600 ngl=100, loglik+grad",
            "eval_unit": "one loglik+grad over a 24x500x2000 trial block", "trials_per_gpu_per_probe": NTRIALS,
            "global_trials_per_probe": NTRIALS * n_gpus, "parallelism": "trial-shard x%d, 1 allreduce of the raw result vector (~70 f64) per eval; the 2 probes run concurrently on 2 host threads / streams; their GEMM phases alternate under the library's token (DESIGN.md 4.1)" % n_gpus,
            "cache": "working set per step 2 x (Y+Z+Zf+B) = 1.5 GB >> 126 MB L2 (inputs larger than L2)",
            "library_env": {"GPCSD_GEMM_RESERVE": os.environ.get("GPCSD_GEMM_RESERVE", "0")}}


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def check_sharded_parity(models, thetas, streams, device, world, torch, dist):
    """N > 1, during warm-up: the trial-sharded evaluation (one all-reduce of the raw result vector) must equal the sum over
    ranks of UNSHARDED evaluations of every rank's own slab -- loglik and its gradient are sums over trials plus a log-det
    term proportional to the trial count (gpcsd1d.py:122-128), so the per-rank values add up exactly.  The per-rank values are
    all-gathered and summed in rank order on the host.  Asserts 1e-12 (loglik) / 1e-11 (gradient, floor 1e-6 of the largest
    component); returns the measured figures for the JSON line."""
    from gpcsd_b200.engine import KronEngine
    out = {"loglik_rel": 0.0, "grad_rel": 0.0}
    for p, m in enumerate(models):
        with torch.cuda.stream(streams[p]):
            m._set_tparams(thetas[p][0], False)
            hp = m._hyperparams()
            eng = m._get_engine()
            ll, g = eng.loglik_grad(hp)                                   # sharded: all-reduced over the ranks
            loc = KronEngine(1, m.x, m.t, m._quadrature(), group=None, jitter=m.JITTER)
            loc.share_data_with(eng)
            loc.ntrials_total = loc.ntrials                                # this rank's slab as a model of its own
            ll_l, g_l = loc.loglik_grad(hp)
            v = torch.tensor([ll_l] + list(g_l), dtype=torch.float64, device=device)
            parts = [torch.zeros_like(v) for _ in range(world)]
            dist.all_gather(parts, v)
            tot = np.zeros(v.numel())
            for q in parts:                                                # fixed rank order
                tot += q.cpu().numpy()
            del loc
        rel_ll = abs(ll - tot[0]) / abs(tot[0])
        rel_g = float(np.max(np.abs(g - tot[1:]) / np.maximum(np.abs(tot[1:]), 1e-6 * np.max(np.abs(tot[1:])))))
        out["loglik_rel"], out["grad_rel"] = max(out["loglik_rel"], rel_ll), max(out["grad_rel"], rel_g)
    if not (out["loglik_rel"] < 1e-12 and out["grad_rel"] < 1e-11):
        raise SystemExit("sharded parity FAILED: %r" % (out,))
    torch.cuda.synchronize()
    return out


def secondary_lines(torch, device, K):
    """Driver-visible numbers for the other single-GPU configurations of BASELINE.json (N = 1 only): device-resident
    loglik+grad through the public model classes, CUDA events around K evaluations after warm-up; hyperparameters change every
    evaluation.  LFP is N(0,1) generated on the device (the arithmetic does not depend on the values)."""
    from gpcsd_b200.gpcsd1d import GPCSD1D
    from gpcsd_b200.gpcsd2d import GPCSD2D
    out = []

    def run(model, tp0, nsteps, label, extra):
        rng = np.random.default_rng(99)
        seq = [tp0 + 0.05 * rng.standard_normal(tp0.shape) for _ in range(nsteps + 3)]
        for k in range(3):
            model.obj_fun_and_grad(seq[k])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(nsteps):
            model.obj_fun_and_grad(seq[3 + k])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / nsteps
        d = {"workload": label, "metric": METRIC, "value": 1e3 / ms, "unit": UNIT, "ms_per_eval": ms, "steps": nsteps}
        d.update(extra)
        out.append(d)

    # configs[0]: GPCSD1D 24 x 50 x 50 (sim_from_gp_1D.py shapes and true parameters), scalar noise, P = 7
    np.random.seed(3)
    x = np.linspace(0.0, 2300.0, 24)[:, None]
    t = np.linspace(0.0, 50.0, 50)[:, None]
    lfp = torch.randn((24, 50, 50), dtype=torch.float64, device=device).cpu().numpy()
    m0 = GPCSD1D(lfp, x, t)
    tp0 = np.log(np.array([100.0 / 100, 200.0 / 100, 20.0, 0.5, 5.0, 0.7, 1e-2]))
    run(m0, tp0, max(50, 10 * K), "BASELINE.json configs[0]: GPCSD1D 24 ch x 50 t x 50 trials, scalar noise (P=7), one model, "
        "evaluations issued one after the other", {"flops_per_eval_algorithmic": 4.0 * 50 * 24 * 50 * (24 + 50)})
    # configs[3]: 64 multi-start restarts of the configs[0] model advanced together: ONE native call evaluates all 64
    # hyperparameter vectors (gpcsd_plan_loglik_grad, restart-batched kernels, CUDA-graph replay)
    from gpcsd_b200.engine import HyperParams
    eng0 = m0._get_engine()
    rng = np.random.default_rng(5)

    def hp_of(tp):
        v = np.exp(tp)
        return HyperParams(R=100.0 * v[0], ells=(100.0 * v[1],), temporal=[(0, v[2], v[3]), (1, v[4], v[5])], sig2n=float(v[6]))
    nb = max(20, 2 * K)
    batches = [[hp_of(tp0 + 0.3 * rng.standard_normal(tp0.shape)) for _ in range(64)] for _ in range(nb + 3)]
    for k in range(3):
        eng0.loglik_grad_batch(batches[k])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(nb):
        eng0.loglik_grad_batch(batches[3 + k])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nb
    out.append({"workload": "BASELINE.json configs[3]: GPCSD1D 24 ch x 50 t x 50 trials, 64 multi-start restarts evaluated in ONE "
                            "restart-batched native call per step (new hyperparameters every step)", "metric": METRIC,
                "value": 64.0 * 1e3 / ms, "unit": UNIT, "ms_per_64_restart_batch": ms, "steps": nb})
    # configs[2]: GPCSD2D Neuropixels-shaped 384 ch (4 x 192 checkerboard) x 250 t x 500 trials, ngl 30 x 120, eps = 1, P = 8
    ch = np.arange(384)
    X = np.stack([np.array([16.0, 48.0, 0.0, 32.0])[ch % 4], 20.0 * np.floor(ch / 2)], axis=1)
    t2 = (0.4 * np.arange(250, dtype=np.float64))[:, None]
    lfp2 = torch.randn((384, 250, 500), dtype=torch.float64, device=device).cpu().numpy()
    m2 = GPCSD2D(lfp2, X, t2, a1=-16.0, b1=64.0, a2=-100.0, b2=float(X[:, 1].max()) + 100.0, ngl1=30, ngl2=120, eps=1.0)
    tp2 = np.log(np.array([100.0 / 100, 40.0 / 100, 200.0 / 100, 5.0, 0.5 / 300.0, 1.0, 0.7 / 300.0, 0.5]))
    run(m2, tp2, max(5, K), "BASELINE.json configs[2]: GPCSD2D Neuropixels-shaped 384 ch x 250 t x 500 trials, ngl 30x120, eps=1, "
        "scalar noise (P=8)", {"flops_per_eval_algorithmic": 4.0 * 500 * 384 * 250 * (384 + 250)})
    out[-1]["tflops_algorithmic"] = out[-1]["flops_per_eval_algorithmic"] / (out[-1]["ms_per_eval"] * 1e-3) * 1e-12
    del m2, lfp2
    torch.cuda.empty_cache()
    # configs[4]: one point of the GPCSD2D scaling sweep (the whole sweep: scripts/sweep_2d.py, profiles/r02_sweep_2d.md):
    # 384 ch x 1000 t x 1000 trials -- temporal halves of order 500 take the cuSOLVER syevd branch
    t4 = (0.4 * np.arange(1000, dtype=np.float64))[:, None]
    lfp4 = torch.randn((384, 1000, 1000), dtype=torch.float64, device=device).cpu().numpy()
    m4 = GPCSD2D(lfp4, X, t4, a1=-16.0, b1=64.0, a2=-100.0, b2=float(X[:, 1].max()) + 100.0, ngl1=30, ngl2=120, eps=1.0)
    run(m4, tp2, 3, "BASELINE.json configs[4] sample point: GPCSD2D 384 ch x 1000 t x 1000 trials, ngl 30x120 (sweep: "
        "profiles/r02_sweep_2d.md)", {"flops_per_eval_algorithmic": 4.0 * 1000 * 384 * 1000 * (384 + 1000)})
    out[-1]["tflops_algorithmic"] = out[-1]["flops_per_eval_algorithmic"] / (out[-1]["ms_per_eval"] * 1e-3) * 1e-12
    del m0, m4, lfp4
    torch.cuda.empty_cache()
    return out


def make_models(device, world, torch, groups=None):
    """Two GPCSD1D models (one per probe) set up like fit_gpcsd_baseline.py:80-89, with model-matched
    synthetic LFP generated on the device (generator only; not part of the measured path)."""
    from gpcsd_b200.covariances import GPCSD1DSpatialCovSE, GPCSDTemporalCovMatern, GPCSDTemporalCovSE
    from gpcsd_b200.gpcsd1d import GPCSD1D
    from gpcsd_b200.priors import GPCSDHalfNormalPrior
    x, t = geometry()
    rank = int(os.environ.get("RANK", "0"))
    from gpcsd_b200.parallel import CollectiveOrder
    order = CollectiveOrder(NPROBES)
    models = []
    for probe in range(NPROBES):
        np.random.seed(10 + probe)
        th = true_hyper(probe)
        spatial_cov = GPCSD1DSpatialCovSE(x, a=A_LO, b=B_HI, ngl=NGL)
        se, mat = GPCSDTemporalCovSE(t), GPCSDTemporalCovMatern(t)
        se.params['ell']['prior'].set_params(30.0, 100.0)
        mat.params['ell']['prior'].set_params(1.0, 20.0)
        sig_pri = [GPCSDHalfNormalPrior(0.1) for _ in range(NX)]
        placeholder = np.zeros((NX, NT, 1))
        m = GPCSD1D(placeholder, x, t, a=A_LO, b=B_HI, ngl=NGL, spatial_cov=spatial_cov, temporal_cov_list=[se, mat],
                    sig2n_prior=sig_pri, distributed=(groups[probe] if world > 1 else False))
        m.lfp_is_local = world > 1
        if world > 1:
            # the probes are evaluated from two host threads on two communicators: keep their collectives in ONE global
            # order on every rank (probe 0, probe 1, probe 0, ...), otherwise two ranks can enqueue them crosswise
            m._collective_order = (order, probe)
        m.R['value'] = th["R"]
        spatial_cov.params['ell']['value'] = th["ell"]
        Ks = spatial_cov.compKphi_1d(th["R"])
        scale = np.trace(Ks) / NX                      # unit-scale LFP
        se.params['ell']['value'], se.params['sigma2']['value'] = th["se"][0], th["se"][1] / scale
        mat.params['ell']['value'], mat.params['sigma2']['value'] = th["matern"][0], th["matern"][1] / scale
        m.sig2n['value'] = th["sig2n"].copy()
        # model-matched draw  Y_r = Ls Z_r Lt^T + sqrt(sig2n) E_r  (torch on the device: data generator only)
        Kt = se.compute_Kt() + mat.compute_Kt()
        g = torch.Generator(device=device)
        g.manual_seed(1000 * (probe + 1) + rank)
        ls, Qs = np.linalg.eigh(Ks + 1e-8 * np.eye(NX))
        lt, Qt = np.linalg.eigh(Kt)
        Ls = torch.from_numpy(Qs * np.sqrt(np.maximum(ls, 0))).to(device)
        Lt = torch.from_numpy(Qt * np.sqrt(np.maximum(lt, 0))).to(device)
        Z = torch.randn((NX, NT, NTRIALS), dtype=torch.float64, device=device, generator=g)
        Y = torch.einsum("ia,ajr->ijr", Ls, Z)
        Y = torch.einsum("ijr,bj->ibr", Y, Lt)
        Y += float(np.sqrt(np.mean(th["sig2n"]))) * torch.randn(Y.shape, dtype=torch.float64, device=device, generator=g)
        host = torch.empty((NX, NT, NTRIALS), dtype=torch.float64).pin_memory()
        host.copy_(Y)
        del Z, Y
        m.lfp = host.numpy()                           # numpy view of PINNED memory
        m._invalidate_lfp()
        models.append(m)
    return models


def theta_sequence(model, n, seed):
    """Log-space hyperparameter vectors theta_true + 0.1 N(0,1), one per step."""
    rng = np.random.default_rng(seed)
    vals = [model.R['value'] / 100.0, model.spatial_cov.params['ell']['value'] / 100.0]
    for tc in model.temporal_cov_list:
        vals += [tc.params['ell']['value'], tc.params['sigma2']['value']]
    vals += list(model.sig2n['value'])
    t0 = np.log(np.array(vals, dtype=np.float64))
    return [t0 + 0.1 * rng.standard_normal(t0.shape) for _ in range(n)]


def measure_fp64_peak(torch, device):
    """cuBLAS DGEMM 4096^3, best of 5 (MEASURED_PEAKS.json carries HBM and bf16 only)."""
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    best = 1e9
    for i in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) * 1e-12


def run_gpu(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=180))
    import __graft_entry__ as ge
    ge.ensure_built()

    # one communicator per probe: the two probes' evaluations run on separate host threads, and collectives
    # issued from different threads on ONE communicator could be ordered differently on different ranks
    groups = [dist.new_group(ranks=list(range(world))) for _ in range(NPROBES)] if world > 1 else None
    models = make_models(device, world, torch, groups)
    K, W = args.steps, max(args.warmup, 3)
    thetas = [theta_sequence(m, 2 * (K + W), 500 + p) for p, m in enumerate(models)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nsteps, offset):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(nsteps):
            fn(offset + s)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    last = {}
    isolate = [False]
    stagger_s = [0.0]

    # The two probes are independent models: evaluate them concurrently, one host thread + CUDA stream each,
    # so one probe's latency-bound eigensolve (16 SMs) overlaps the other's DMMA GEMMs (--serial disables this).
    from concurrent.futures import ThreadPoolExecutor
    streams = [torch.cuda.Stream(device=device) for _ in models]
    pool = ThreadPoolExecutor(max_workers=len(models))

    def eval_probe(p, s, upload):
        torch.cuda.set_device(device)
        with torch.cuda.stream(streams[p]):
            m = models[p]
            if upload:
                m.update_lfp(m.lfp, m.t)                   # forces the host->device copy of the pinned block
            return m.obj_fun_and_grad(thetas[p][s])

    def run_steps(first, nsteps, upload):
        """nsteps evaluations of every probe.  Serial mode: probes one after the other inside each step.
        Concurrent mode: one free-running host thread per probe (no per-step join), so in steady state one
        probe's host->device upload / eigensolve overlaps the other probe's GEMMs."""
        if args.serial:
            for s in range(first, first + nsteps):
                for p in range(len(models)):
                    last[p] = eval_probe(p, s, upload)
                    if isolate[0]:
                        torch.cuda.synchronize()           # per-kernel timing pass: nothing of the next evaluation overlaps
            return

        def worker(p):
            # Optional phase offset between the probes (--stagger; round 1 needed it to keep the probes out of lock step).
            # The library now alternates the probes' GEMM phases itself (gpcsd_plan_loglik_grad's token), so the default is
            # to start together.  When used, the offset is inside the timed region.
            if p > 0 and stagger_s[0] > 0.0:
                time.sleep(p * stagger_s[0] / len(models))
            r = None
            for s in range(first, first + nsteps):
                r = eval_probe(p, s, upload)
            return r
        order = getattr(models[0], "_collective_order", (None,))[0]
        if order is not None:
            order.start()                                  # concurrent phase: collectives in one global rotation
        futs = [pool.submit(worker, p) for p in range(len(models))]
        try:
            for p, f in enumerate(futs):
                last[p] = f.result()
        finally:
            if order is not None:
                order.stop()

    def timed_steps(nsteps, first, upload):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        run_steps(first, nsteps, upload)
        for st_ in streams:
            torch.cuda.current_stream(device).wait_stream(st_)
        e1.record()
        barrier()
        ms = max(e0.elapsed_time(e1), 0.0)
        ms = max(ms, 1e3 * 0.0)
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    def step_resident(s):
        run_steps(s, 1, False)

    # ---- device-resident arm ("value")
    for p, m in enumerate(models):
        with torch.cuda.stream(streams[p]):
            m._get_engine()                                 # upload once
    torch.cuda.synchronize()

    def serial_once(first, upload):
        """First use of every code path strictly serial: CUDA loads kernels lazily on their first launch and workspaces,
        streams and events are created on first use -- all of which may need the device to drain, which deadlocks if the
        other probe's thread already sits in a blocking NCCL all-reduce whose peer rank does the same thing the other way
        round.  After this step the concurrent phases launch nothing new."""
        prev = args.serial
        args.serial = True
        run_steps(first, 1, upload)
        torch.cuda.synchronize()
        args.serial = prev

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()         # nvidia-smi needs ~0.1 s to start streaming: begin before the warm-up, all of it is under load
    serial_once(0, False)
    sharded_parity = None
    if world > 1:
        sharded_parity = check_sharded_parity(models, thetas, streams, device, world, torch, dist)
    if not args.no_stagger:
        # one evaluation of one probe, wall clock (device drained before and after): the phase offset is half of it
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eval_probe(0, 0, False)
        torch.cuda.synchronize()
        t_eval = time.perf_counter() - t0
        if world > 1:
            # every rank uses the same offset; the slowest rank's estimate
            tt = torch.tensor([t_eval], dtype=torch.float64, device=device)
            eval_probe(1, 0, False)                        # keep the probes' collective counts equal
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_eval = float(tt.item())
        stagger_s[0] = t_eval * float(os.environ.get("GPCSD_BENCH_STAGGER", "1.0"))
    timed_steps(W, 0, False)
    if args.profile_step:
        # one steady-state SERIAL step between cudaProfilerStart/Stop for `ncu --profile-from-start off`
        args.serial = True
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident(W)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if rank == 0:
            sampler.stop()
            print(json.dumps({"profile_step": "done"}))
        return
    engines = [m._get_engine() for m in models]
    # pass 1 (headline): the two probes evaluated concurrently
    for e in engines:
        e.n_launches = 0
    ms_total = timed_steps(K, W, False)
    launches = sum(e.n_launches for e in engines)
    # pass 2: same K steps with the probes one after the other and CUDA events around the dominant kernels
    # (per-kernel durations are only meaningful without a second stream competing for the SMs)
    concurrent = not args.serial
    args.serial = True
    timed_steps(2, 0, False)
    isolate[0] = True
    ms_serial = timed_steps(K, W, False)                    # product path (native plan), one evaluation at a time
    # per-kernel durations: the same kernels launched call by call from Python (engine's stepwise path) with CUDA events on
    # the launching stream around every ABI call; first warm that path (its workspaces are allocated on first use)
    names = ["gpcsd_project_quad", "gpcsd_project_quad_strided", "gpcsd_wsyrk", "gpcsd_eigh", "gpcsd_eigh_dc", "gpcsd_dgemm"]
    for e in engines:
        e.timers = {n: [] for n in names}
    timed_steps(2, 0, False)
    for e in engines:
        e.timers = {n: [] for n in names}
    timed_steps(K, W, False)
    isolate[0] = False
    kt = {}
    for name in engines[0].timers:
        d = [a.elapsed_time(b) for e in engines for (a, b) in e.timers[name]]
        kt[name] = (float(np.mean(d)) if d else 0.0, len(d))
    for e in engines:
        e.timers = None
    args.serial = not concurrent
    # ---- end-to-end arm
    serial_once(K + W, True)
    timed_steps(W, K + W, True)
    ms_e2e = timed_steps(K, K + 2 * W, True)
    clocks = sampler.stop() if rank == 0 else None      # sampled every 20 ms across the three timed passes above

    # ---- second half of the metric: posterior CSD prediction (type="csd", z = electrode sites), trials/s
    KP = max(2, K // 3)

    def predict_probe(p, to_host):
        torch.cuda.set_device(device)
        with torch.cuda.stream(streams[p]):
            m = models[p]
            if to_host:
                m.predict(m.x, m.t, type="csd")          # public API: results land in host numpy arrays
                return float(m.csd_pred[0, 0, 0])
            out = m._get_engine().predict(m._hyperparams(), m.x, m.t, kind="csd", to_host=False)
            return out["csd_pred"]

    def step_predict(to_host):
        def run(_s):
            futs = [pool.submit(predict_probe, p, to_host) for p in range(len(models))]
            return [f.result() for f in futs]
        return run

    for m, th in zip(models, thetas):
        m._set_tparams(th[0], False)
    for to_host in (False, True):                          # first use serial (see serial_once)
        for p in range(len(models)):
            predict_probe(p, to_host)
        torch.cuda.synchronize()
    # device-resident predict: free-running worker per probe with the same phase offset as the evaluations
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    predict_probe(0, False)
    torch.cuda.synchronize()
    t_pred = time.perf_counter() - t0          # (predict does not go through the plan's token: the phase offset stays)

    def timed_predict(nsteps):
        def worker(p):
            if p > 0 and t_pred > 0.0:
                time.sleep(p * t_pred / len(models))
            r = None
            for _ in range(nsteps):
                r = predict_probe(p, False)
            return r
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        futs = [pool.submit(worker, p) for p in range(len(models))]
        for f in futs:
            f.result()
        for st_ in streams:
            torch.cuda.current_stream(device).wait_stream(st_)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    KPR = max(KP, K)
    timed_predict(2)
    ms_pred = timed_predict(KPR) * KP / KPR             # normalised to KP steps like the end-to-end arm below
    timed(step_predict(True), 3, 0)                        # lets torch's pinned-host cache reach steady state
    ms_pred_e2e = timed(step_predict(True), KP, 0)
    trials_per_step = NPROBES * NTRIALS * world
    ntc = 2
    predict_line = {"metric": "gpcsd_csd_predict_trials_per_s", "unit": "trials/s",
                    "value": trials_per_step * KP / (ms_pred * 1e-3), "ms_per_step": ms_pred / KP,
                    "e2e": {"value": trials_per_step * KP / (ms_pred_e2e * 1e-3), "ms_per_step": ms_pred_e2e / KP,
                            "d2h_bytes_per_step": NPROBES * (ntc + 1) * NX * NT * NTRIALS * 8,
                            "api": "GPCSD1D.predict(x, t, type='csd') -> csd_pred + csd_pred_list as host arrays"},
                    "steps": KP, "config": "z = the 24 electrode sites, t* = t, per-component (SE, Matern) + summed CSD"}

    # predict roofline: the temporal back-projection GEMM (two half-order blocks per temporal component and probe), timed
    # with CUDA events in a serial pass (one probe at a time, device drained in between)
    for e in engines:
        e.timers = {"predict_backproject": []}
    for _ in range(2):
        for p in range(len(models)):
            predict_probe(p, False)
            torch.cuda.synchronize()
    d = [a.elapsed_time(b) for e in engines for (a, b) in e.timers["predict_backproject"]]
    for e in engines:
        e.timers = None
    pred_kernel_ms = float(np.mean(d)) if d else None
    pred_flops = 2.0 * (NT // 2) ** 2 * NTRIALS * NX            # one block launch: nz = NX batched (NT/2)^2 x trials products

    # strong scaling (what configs[1] literally says: 2000 trials split over the N GPUs): every rank keeps 2000/N trials
    strong = None
    if world > 1:
        nloc = NTRIALS // world
        for m in models:
            m.lfp = np.ascontiguousarray(m.lfp[:, :, :nloc])
            m._invalidate_lfp()
        for p, m in enumerate(models):
            with torch.cuda.stream(streams[p]):
                m._get_engine()
        torch.cuda.synchronize()
        serial_once(0, False)
        timed_steps(W, 0, False)
        ms_strong = timed_steps(K, W, False)
        strong = {"scaling": "strong", "value": NPROBES * K / (ms_strong * 1e-3), "unit": "evals/s of the full 2000-trial model",
                  "ms_per_step": ms_strong / K, "trials_per_gpu_per_probe": nloc, "global_trials_per_probe": nloc * world,
                  "amdahl_term": "covariance build + the two eigendecompositions are replicated on every rank and do not shrink "
                                 "with N (gpcsd_eigh_dc %.3f ms per evaluation, see kernel_ms); the trial-proportional GEMM/SYRK "
                                 "work divides by N" % kt["gpcsd_eigh_dc"][0]}

    evals_per_step = NPROBES * world
    value = evals_per_step * K / (ms_total * 1e-3)
    e2e_value = evals_per_step * K / (ms_e2e * 1e-3)
    P = 6 + NX
    if rank == 0:
        peak = measure_fp64_peak(torch, device)
        # projection Qt^T Z_i for all i.  On the uniform time grid of this workload Qt is block diagonal in
        # the folded time basis, so the projection is two launches of order NT/2 -- half the flops of the reference's
        # contraction; the roofline line is per launch of the dominant kernel as executed.
        folded = kt["gpcsd_project_quad_strided"][1] > 0
        mblk = NT // 2 if folded else NT
        algo_flops = 2.0 * NX * mblk * mblk * NTRIALS
        dur_ms = kt["gpcsd_project_quad_strided" if folded else "gpcsd_project_quad"][0]
        achieved = algo_flops / (dur_ms * 1e-3) * 1e-12 if dur_ms > 0 else None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("project_quad_dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "value_serial": evals_per_step * K / (ms_serial * 1e-3),
                "ms_per_step_serial": ms_serial / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(world),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / K,
                        "h2d_bytes_per_step": NPROBES * NX * NT * NTRIALS * 8 + NPROBES * 0,
                        "d2h_bytes_per_step": NPROBES * (8 + 2 * 2 + 2 * NX) * 8,
                        "api": "GPCSD1D.update_lfp(pinned lfp) + GPCSD1D.obj_fun_and_grad(tparams)"},
                "gpu_launches": launches,
                "roofline": {"bound": "tensor", "kernel": "tma_gemm_kernel<NN,EPI_QUAD,NTW=3> (gpcsd_project_quad_strided, one of the two half-order blocks of the folded time basis: persistent TMA+mbarrier DMMA GEMM, 128x96 tiles, fused /D + quadratic form)",
                             "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                             "algorithmic_flops_per_launch": algo_flops, "avg_launch_ms": dur_ms,
                             "timed_in": "serial pass (value_serial: one evaluation at a time, device drained in between): CUDA events on the launching stream around every ABI call of the kernel (the span also holds its 1-CTA partial-sum reduction)",
                             "peak_source": "in-run cuBLAS DGEMM 4096^3 best-of-5 (FP64; MEASURED_PEAKS.json has only "
                                            "HBM and bf16); DMMA issue-rate microbenchmark: 37.0 TFLOP/s"},
                "kernel_ms": {k: {"avg_ms": v[0], "calls": v[1]} for k, v in kt.items()},
                "last_nll": [float(last[p][0]) for p in range(NPROBES)],
                "predict": predict_line}
        if pred_kernel_ms:
            ach = pred_flops / (pred_kernel_ms * 1e-3) * 1e-12
            line["predict"]["roofline"] = {
                "bound": "tensor", "kernel": "tma_gemm_kernel<NN,EPI_STORE> (gpcsd_dgemm: temporal back-projection (C_k U)(V_z) of one "
                                             "half-order block of the folded time basis, batched over the nz = 24 output sites)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "algorithmic_flops_per_launch": pred_flops, "avg_launch_ms": pred_kernel_ms,
                "peak_source": "in-run cuBLAS DGEMM 4096^3 (FP64)"}
        if sharded_parity is not None:
            line["sharded_parity_rel"] = sharded_parity
        if strong is not None:
            line["strong"] = strong
        if world == 1 and not args.no_secondary:
            line["secondary"] = secondary_lines(torch, device, K)
        if world == 1 and not args.no_cpu_baseline:
            nthr, blas = blas_threads()
            cpu_eval_seconds(reps=1, seed0=90)                      # warm the BLAS threads
            secs = cpu_eval_seconds(reps=5)
            line["cpu_baseline"] = {"value": 1.0 / float(np.median(secs)), "unit": UNIT, "cores": nthr, "host_cpus": os.cpu_count(),
                                    "blas": blas, "kind": "port",
                                    "sample": "oracle loglik_and_grad (numpy/LAPACK restatement; closed-form gradient) on one FULL "
                                              "24x500x2000 block, median of 5 evaluations, nothing extrapolated"}
            line["predict"]["cpu_baseline"] = cpu_predict_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[0]/[2] secondary lines (N = 1)")
    ap.add_argument("--serial", action="store_true", help="evaluate the two probes one after the other")
    ap.add_argument("--stagger", dest="no_stagger", action="store_false",
                    help="start the probes' evaluation loops half an evaluation apart (round-1 behaviour; the library's GEMM token "
                         "now alternates the probes by itself: 850 evals/s without the offset, 812-855 with it)")
    ap.add_argument("--no-stagger", dest="no_stagger", action="store_true", help="(default) start the probes' evaluation loops together")
    ap.set_defaults(no_stagger=True)
    ap.add_argument("--profile-step", action="store_true", help="run warm-up then ONE step inside cudaProfilerStart/Stop")
    args = ap.parse_args()
    # watchdog: a hang (e.g. a collective one rank never enters) becomes a stack dump of every thread and a non-zero exit
    # instead of a silent stall that holds the GPU box until the caller's timeout
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ.get("GPCSD_BENCH_WATCHDOG_S", "600")), exit=True)
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
        return
    run_gpu(args)


if __name__ == "__main__":
    main()
