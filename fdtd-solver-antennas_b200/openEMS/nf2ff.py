"""openEMS.nf2ff stand-in: the object returned by FDTD.CreateNF2FFBox() and its CalcNF2FF
(antenna_sim/solver_fdtd_openems_microstrip_3d.py:179,225; App. A6)."""
from __future__ import annotations

import numpy as np

from b200fdtd import postproc


class nf2ff_results:
    """Result container with openEMS's attribute names (fields are lists indexed by frequency)."""

    def __init__(self, freq, theta, phi, r, parts):
        self.freq = np.atleast_1d(np.asarray(freq, np.float64))
        self.theta = np.asarray(theta, np.float64)
        self.phi = np.asarray(phi, np.float64)
        self.r = float(r)
        self.Dmax = np.array([p["Dmax"] for p in parts])
        self.Prad = np.array([p["Prad"] for p in parts])
        self.E_theta = [p["E_theta"] for p in parts]
        self.E_phi = [p["E_phi"] for p in parts]
        self.E_norm = [np.sqrt(np.abs(p["E_theta"]) ** 2 + np.abs(p["E_phi"]) ** 2) for p in parts]
        self.E_cprh, self.E_cplh = [], []
        for p in parts:
            c, s = np.cos(np.deg2rad(self.phi))[None, :], np.sin(np.deg2rad(self.phi))[None, :]
            self.E_cprh.append((c + 1j * s) * (p["E_theta"] + 1j * p["E_phi"]) / np.sqrt(2.0))
            self.E_cplh.append((c - 1j * s) * (p["E_theta"] - 1j * p["E_phi"]) / np.sqrt(2.0))
        self.P_rad = [p["P_rad"] for p in parts]


class nf2ff:
    def __init__(self, CSX, name, start, stop, **kw):
        self.CSX = CSX
        self.name = name
        self.start = np.asarray(start, np.float64)
        self.stop = np.asarray(stop, np.float64)
        self.freq = kw.get("frequency", None)
        self.dump_type = 10 if self.freq is not None else 0
        self._fdtd = None
        add_line = kw.get("add_mesh_line", False)
        if add_line:
            g = CSX.GetGrid()
            for a in range(3):
                g.AddLine(a, [self.start[a], self.stop[a]])

    def CalcNF2FF(self, sim_path, freq, theta, phi, radius=1, center=[0, 0, 0], outfile=None, read_cached=False, verbose=0):
        """theta/phi in degrees; center in metres; returns nf2ff_results"""
        from . import _registry as _o
        res = _o.results_for(sim_path, self._fdtd)
        if res is None or "nf2ff" not in res:
            raise RuntimeError(f"CalcNF2FF: no finished simulation with an NF2FF box found for '{sim_path}'")
        freqs = np.atleast_1d(np.asarray(freq, np.float64))
        theta = np.atleast_1d(np.asarray(theta, np.float64)); phi = np.atleast_1d(np.asarray(phi, np.float64))
        dev = res.get("device", 0)
        parts = [postproc.far_field(res["nf2ff"], f, theta, phi, center=center, radius=float(radius),
                                    farfield_fn=res.get("farfield_fn"), device=dev) for f in freqs]
        return nf2ff_results(freqs, theta, phi, radius, parts)
