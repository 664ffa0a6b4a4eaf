"""-m gpu: the native evaluation plan (gpcsd_plan_*: one C-ABI call per evaluation, restart-batched, CUDA-graph replay)
against the oracle and against the call-by-call Python orchestration of the same kernels."""
import numpy as np
import pytest

from helpers import engine_from_oracle, hp_from_oracle, scaled_rel, ulp_sensitivity

pytestmark = pytest.mark.gpu


def _thetas(om, n, seed, scale=0.15):
    from oracle import gpcsd_oracle as O
    rng = np.random.default_rng(seed)
    tp = O.pack_tparams(om)
    return [O.unpack_tparams(om, tp + scale * rng.standard_normal(tp.shape)) for _ in range(n)]


def test_64_restart_batch_matches_oracle_config0_shape(cuda_lib):
    """BASELINE configs[0]/[3] shape: 24 x 50 x 50, 64 hyperparameter vectors in ONE call, each at 1e-9 vs the oracle."""
    from oracle import gpcsd_oracle as O, synth
    x, t = synth.geometry_1d(24, 50)
    om = synth.model_1d(x, t, sig2n=1e-2)
    lfp = synth.matched_lfp(om, 50, 1000)
    eng, _ = engine_from_oracle(om, lfp)
    oms = _thetas(om, 64, 4)
    hps = [hp_from_oracle(m) for m in oms]
    for rep in range(3):                                  # eager, capture, replay: identical results
        ll, g, flag = eng.loglik_grad_batch(hps)
        assert np.all(flag == 0)
        if rep == 0:
            ll0, g0 = ll.copy(), g.copy()
        else:
            assert np.array_equal(ll, ll0) and np.array_equal(g, g0)
    worst_ll = worst_g = 0.0
    for r, m in enumerate(oms):
        ll_o, g_o = O.loglik_and_grad(m, lfp)
        worst_ll = max(worst_ll, abs(ll[r] - ll_o) / abs(ll_o))
        worst_g = max(worst_g, float(np.max(np.abs(g[r] - g_o) / np.abs(g_o))))
    print("\n[plan, 64 restarts 24x50x50] worst loglik rel %.2e, worst grad rel %.2e" % (worst_ll, worst_g))
    assert worst_ll < 1e-9 and worst_g < 1e-9
    # a batch equals the same evaluations issued one by one, bit for bit (same kernels, same order per restart)
    ll1, g1 = eng.loglik_grad(hps[5])
    assert abs(ll1 - ll[5]) <= 1e-13 * abs(ll1) and np.max(np.abs(g1 - g[5]) / np.abs(g1)) < 1e-12


@pytest.mark.parametrize("case", ["vec_nt130_fold", "scalar_nt64", "odd_nt41", "nonuniform_nt40", "grid2d", "checkerboard2d"])
def test_plan_equals_stepwise_path_and_oracle(cuda_lib, case):
    from gpcsd_b200.engine import KronEngine
    from oracle import gpcsd_oracle as O, synth
    N = 9
    if case == "vec_nt130_fold":
        x, t = synth.geometry_1d(24, 130)
        om = synth.model_1d(x, t, sig2n=1e-2 * np.exp(0.3 * np.random.default_rng(0).standard_normal(24)))
    elif case == "scalar_nt64":
        x, t = synth.geometry_1d(24, 64)
        om = synth.model_1d(x, t, sig2n=1e-3)
    elif case == "odd_nt41":
        x, t = synth.geometry_1d(24, 41)
        om = synth.model_1d(x, t)
    elif case == "nonuniform_nt40":
        x, t = synth.geometry_1d(24, 40)
        om = synth.model_1d(x, t + 0.05 * np.sin(np.arange(40))[:, None])
    elif case == "grid2d":
        X, t = synth.geometry_grid_2d(4, 12, 30)
        om = synth.model_2d(X, t, ngl1=8, ngl2=24, sig2n=0.3)
    else:
        X, t = synth.geometry_neuropixels(96, 40, 0.4)
        om = synth.model_2d(X, t, ngl1=10, ngl2=40, a1=-16.0, b1=64.0, a2=-100.0, b2=float(X[:, 1].max()) + 100.0, eps=1.0, sig2n=0.5)
    lfp = synth.matched_lfp(om, N, 21)
    oms = _thetas(om, 3, 7, scale=0.05)
    hps = [hp_from_oracle(m) for m in oms]
    eng, _ = engine_from_oracle(om, lfp)
    assert eng.use_plan
    ll, g, flag = eng.loglik_grad_batch(hps)
    ll2, g2, _ = eng.loglik_grad_batch(hps)               # second call records the graph, third replays it
    ll3, g3, _ = eng.loglik_grad_batch(hps)
    assert np.array_equal(ll, ll2) and np.array_equal(ll, ll3) and np.array_equal(g, g3)
    step, _ = engine_from_oracle(om, lfp)
    step.use_plan = False
    vec = np.ndim(om.sig2n) > 0
    for r, m in enumerate(oms):
        ll_s, g_s = step.loglik_grad(hps[r])
        ll_o, g_o = O.loglik_and_grad(m, lfp)
        # per-electrode noise: the reference formula moves by `sens` under 1-ulp perturbations of (Ks, Kt); a backward-
        # stable eigensolver guarantees a backward error of O(n) ulps (n = 24), so n x sens is the agreement two correct
        # implementations can be held to on a 9-trial sample (the full-size tests use 4 x sens and pass at 0.1-0.7 x sens)
        tol = max(1e-9, 24.0 * ulp_sensitivity(m, lfp, nseeds=2)) if vec else 1e-9
        assert abs(ll[r] - ll_o) / abs(ll_o) < 1e-9
        assert scaled_rel(g[r], g_o).max() < tol, (case, r)
        assert abs(ll[r] - ll_s) <= 1e-12 * abs(ll_s)
        assert scaled_rel(g[r], g_s).max() < max(1e-11, tol)
        assert abs(eng.loglik(hps[r]) - ll_o) / abs(ll_o) < 1e-9


def test_plan_eigensolver_failure_flag(cuda_lib):
    """NaN hyperparameters: numpy's eigh raises LinAlgError; the plan reports it per restart and the engine raises."""
    from oracle import synth
    x, t = synth.geometry_1d(24, 40)
    om = synth.model_1d(x, t)
    lfp = synth.matched_lfp(om, 5, 2)
    eng, hp = engine_from_oracle(om, lfp)
    import copy
    bad = copy.deepcopy(hp)
    bad.R = float("nan")
    ll, g, flag = eng.loglik_grad_batch([hp, bad, hp])
    assert flag[0] == 0 and flag[2] == 0 and flag[1] != 0
    assert np.isfinite(ll[0]) and ll[0] == ll[2]
    with pytest.raises(np.linalg.LinAlgError):
        eng.loglik_grad(bad)


_TWO_PHASE_SCRIPT = r"""
import json, os, sys
sys.path.insert(0, os.path.join(%(root)r, "tests")); sys.path.insert(0, %(root)r)
import numpy as np
from helpers import engine_from_oracle, hp_from_oracle
from oracle import gpcsd_oracle as O, synth
x, t = synth.geometry_1d(24, 130)
om = synth.model_1d(x, t, sig2n=1e-2 * np.exp(0.3 * np.random.default_rng(0).standard_normal(24)))
lfp = synth.matched_lfp(om, 40, 7)
eng, _ = engine_from_oracle(om, lfp)
rng = np.random.default_rng(3)
tp = O.pack_tparams(om)
hps = [hp_from_oracle(O.unpack_tparams(om, tp + 0.1 * rng.standard_normal(tp.shape))) for _ in range(3)]
out = []
for rep in range(4):                                     # eager, capture, replay of whichever path the policy selects
    ll, g = eng.loglik_grad(hps[0])
    llb, gb, flag = eng.loglik_grad_batch(hps)
    out.append({"ll": float(ll), "g": np.asarray(g).tolist(), "llb": np.asarray(llb).tolist(), "gb": np.asarray(gb).tolist()})
print("RESULT " + json.dumps(out))
"""


def test_two_phase_token_path_matches_single_launch(cuda_lib, tmp_path):
    """gpcsd_plan_loglik_grad issues the evaluation as two launches (prologue | GEMM phase under the process-wide token) when
    calls from several host threads overlap; GPCSD_GEMM_TOKEN=1 forces that path.  Same kernels, same grids: bit-identical to
    the single-launch path.  With GPCSD_GEMM_RESERVE=16 the GEMM phase is sized for all SMs but 16 (partial sums are grouped
    differently): equal to rounding.  Eager / captured / replayed launches of every path agree bit for bit."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "two_phase.py"
    script.write_text(_TWO_PHASE_SCRIPT % {"root": root})
    res = {}
    for mode in ("0", "1", "reserve"):
        env = dict(os.environ, GPCSD_GEMM_TOKEN="0" if mode == "0" else "1", GPCSD_GEMM_RESERVE="16" if mode == "reserve" else "0")
        p = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
        res[mode] = json.loads(line[len("RESULT "):])
    for mode in ("0", "1"):
        for rep in res[mode][1:]:
            assert rep == res[mode][0]                    # eager == captured == replayed, bit for bit
    for rep in res["reserve"][2:]:
        assert rep == res["reserve"][1]                   # (the very first call after an upload keeps the one-launch path)
    assert res["1"][0] == res["0"][0]                     # two launches, same grids: bit-identical
    a, b = res["0"][0], res["reserve"][1]
    assert abs(a["ll"] - b["ll"]) <= 1e-13 * abs(a["ll"])
    ga, gb = np.array(a["g"]), np.array(b["g"])
    # (per-electrode noise: the (R, ell) components amplify rounding ~1e5 x)
    assert np.max(np.abs(ga - gb) / np.maximum(np.abs(ga), 1e-300)) < 1e-8
    assert np.max(np.abs(np.array(a["llb"]) - np.array(b["llb"])) / np.abs(np.array(a["llb"]))) <= 1e-13
    assert np.max(np.abs(np.array(a["gb"]) - np.array(b["gb"])) / np.maximum(np.abs(np.array(a["gb"])), 1e-300)) < 1e-8
