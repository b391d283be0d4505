"""Physical constants with the values of openEMS.physical_constants (imported by the reference at
antenna_sim/solver_fdtd_openems_microstrip_3d.py:41)."""
import math

C0 = 299792458.0
MUE0 = 4e-7 * math.pi
EPS0 = 1.0 / (MUE0 * C0 * C0)
Z0 = math.sqrt(MUE0 / EPS0)
EPS0_ENERGY = 8.85418781762e-12   # constants of the engine's energy estimate (csrc/b200fdtd.cu)
MUE0_ENERGY = 1.256637062e-6
