"""gpcsd_b200 -- B200-native GPCSD hot path (loglik + gradient, predict) behind gpcsd's Python API."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
