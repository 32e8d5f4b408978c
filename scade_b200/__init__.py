"""scade_b200 -- B200-native (sm_100a) renderer for the SCADE NeRF hot path.

Public surface mirrors the reference (mikacuy/scade): see nerf_helpers.py (model/run_nerf_helpers.py)
and render.py (renderer half of run_scade_scannet.py).  Importing the package does not load CUDA;
the first kernel call builds/loads libscade_b200.so and raises if that is impossible.
"""
from . import synthetic  # noqa: F401

__all__ = ["synthetic", "nerf_helpers", "render", "functional", "dist"]
__version__ = "0.1.0"
