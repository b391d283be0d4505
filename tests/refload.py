"""Load the reference's solver modules WITHOUT executing antenna_sim/__init__.py (it imports matplotlib)
and without touching the reference tree.  Only usable where /root/reference exists
(the build container); GPU-box tests use the recorded call traces under tests/golden/ instead."""
import importlib.util
import os
import sys
import types

REF = "/root/reference"
PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fdtd-solver-antennas_b200")
DLL_DIR = os.path.join(PKG, "dll_sentinel")


def available():
    return os.path.isdir(os.path.join(REF, "antenna_sim"))


def load(names=("models", "physics", "solver_fdtd_openems_fixed", "solver_fdtd_openems_microstrip",
                "solver_fdtd_openems_microstrip_3d", "solver_fdtd_openems_microstrip_multi_3d")):
    if not hasattr(os, "add_dll_directory"):
        os.add_dll_directory = lambda p: None
    if "antenna_sim" not in sys.modules or not hasattr(sys.modules["antenna_sim"], "__b200_stub__"):
        pkg = types.ModuleType("antenna_sim")
        pkg.__path__ = [os.path.join(REF, "antenna_sim")]
        pkg.__b200_stub__ = True
        sys.modules["antenna_sim"] = pkg
    mods = {}
    for n in names:
        full = "antenna_sim." + n
        if full not in sys.modules:
            spec = importlib.util.spec_from_file_location(full, os.path.join(REF, "antenna_sim", n + ".py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[full] = m
            spec.loader.exec_module(m)
        mods[n] = sys.modules[full]
    return mods
