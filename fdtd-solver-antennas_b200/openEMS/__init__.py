"""openEMS — drop-in stand-in for the openEMS Python package, backed by the B200 FDTD engine.

  from openEMS import openEMS                          antenna_sim/solver_fdtd_openems_microstrip_3d.py:40
  from openEMS.physical_constants import C0, EPS0      …microstrip_3d.py:41
  from openEMS import CSXCAD, nf2ff, ports, utilities, automesh   antenna_sim/solver_fdtd_openems.py:117-124
"""
import os as _os

# the reference's Windows glue calls os.add_dll_directory (antenna_sim/solver_fdtd_openems_microstrip.py:74-81);
# on Linux the attribute is absent, so provide a no-op
if not hasattr(_os, "add_dll_directory"):
    _os.add_dll_directory = lambda p: None

import CSXCAD  # noqa: E402,F401
from .openEMS import openEMS  # noqa: E402,F401
from . import physical_constants, ports, nf2ff, utilities, automesh  # noqa: E402,F401

__version__ = "0.0.36-b200"
