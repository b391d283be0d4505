"""openEMS.utilities stand-in (legacy probe imports it: antenna_sim/solver_fdtd_openems.py:117-124)."""
import numpy as np

from b200fdtd.postproc import dft_time2freq


def DFT_time2freq(t, val, freq, signal_type="pulse"):
    if signal_type != "pulse":
        raise NotImplementedError("only pulse signals are used by the reference")
    return dft_time2freq(t, val, freq)


def Check_Array_Equal(a, b, tol, relative=False):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    d = np.abs(a - b)
    if relative:
        d = d / np.maximum(np.abs(a), 1e-300)
    return bool(np.all(d < tol))
