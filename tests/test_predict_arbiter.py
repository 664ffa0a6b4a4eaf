"""CPU: which float64 formulation of predict is right when they disagree?  The reference's dense inverse (gpcsd1d.py:263-265)
and the Kronecker form (oracle.predict_kron, and the CUDA engine) against a 40-digit mpmath solve of the same linear system
(oracle/arbiter_mp.py) at cond(K) ~ 1e9: the Kronecker form is the accurate one, which is why the parity gate against the
reference-generated golden predictions is 1e-6 while the gate against predict_kron is 1e-8."""
import numpy as np

from helpers import relerr


def test_kronecker_form_is_the_accurate_one():
    from oracle import gpcsd_oracle as O, synth
    from oracle.arbiter_mp import predict_arbiter
    x, t = synth.geometry_1d(10, 14)
    om = synth.model_1d(x, t, sig2n=1e-7)                       # low noise: K = Ks (x) Kt + sig2n I is ill-conditioned
    lfp = synth.matched_lfp(om, 2, 5)
    z = np.linspace(100.0, 2200.0, 5)[:, None]
    Ks, Kt = om.Ks(), om.Kt()
    cond = (np.linalg.eigvalsh(Ks)[-1] * np.linalg.eigvalsh(Kt)[-1] + 1e-7) / 1e-7
    assert cond > 1e7
    arb = predict_arbiter(om, lfp, z, "csd")
    dense = O.predict_dense(om, lfp, z, om.t, "csd")["csd_pred_list"]
    kron = O.predict_kron(om, lfp, z, om.t, "csd")["csd_pred_list"]
    for c in range(2):
        e_dense, e_kron = relerr(dense[c], arb[c]), relerr(kron[c], arb[c])
        print("\ncomponent %d (cond %.1e): reference dense formulation %.2e, Kronecker form %.2e from the 40-digit result" % (c, cond, e_dense, e_kron))
        assert e_kron < 1e-9
        assert e_kron < 0.1 * e_dense
