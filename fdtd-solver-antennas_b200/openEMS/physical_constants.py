"""openEMS.physical_constants (imported at antenna_sim/solver_fdtd_openems_microstrip_3d.py:41)."""
from b200fdtd.constants import C0, MUE0, EPS0, Z0  # noqa: F401
