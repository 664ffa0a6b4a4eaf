"""Drop-in alias: put ``<repo>/dropin`` on PYTHONPATH and existing scripts that do
``from gpcsd.gpcsd1d import GPCSD1D`` run on the B200 engine unchanged."""
import importlib
import sys

_MODULES = ["covariances", "forward_models", "predict_csd", "priors", "utility_functions", "gpcsd1d", "gpcsd2d"]
for _m in _MODULES:
    sys.modules[__name__ + "." + _m] = importlib.import_module("gpcsd_b200." + _m)
    globals()[_m] = sys.modules[__name__ + "." + _m]
