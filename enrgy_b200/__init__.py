"""enrgy_b200 -- B200-native drop-in for the per-cell, per-timestep surface energy balance of
tepextepex/ENRGY (`Energy.model`, reference model.py:155-286).

`from enrgy_b200 import Energy` gives the reference-shaped class; `enrgy_b200.engine.Engine` is the
thin wrapper over the C ABI (include/enrgy_b200.h).  The CUDA library is loaded lazily on first use
and there is no CPU fallback.
"""
from .model import Energy, PARAMS  # noqa: F401

__all__ = ["Energy", "PARAMS"]
