"""CSXCAD — drop-in stand-in for the CSXCAD Python package, backed by the B200 FDTD engine.

Only the call surface the reference uses is provided:
  from CSXCAD import ContinuousStructure        antenna_sim/solver_fdtd_openems_microstrip_3d.py:39
  CSXCAD.CSProperties / CSXCAD.CSPrimitives     antenna_sim/solver_fdtd_openems_fixed.py:104-105 (probe dir())
"""
from .CSXCAD import ContinuousStructure  # noqa: F401
from . import CSProperties, CSPrimitives, CSRectGrid, CSTransform, SmoothMeshLines  # noqa: F401

__version__ = "0.6.3-b200"
