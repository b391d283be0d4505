"""openEMS.ports stand-in: LumpedPort with the CalcPort contract.

Reference call sites: FDTD.AddLumpedPort(1, 50.0, start, stop, 'z', 1.0, priority=5, edges2grid='xy')
(antenna_sim/solver_fdtd_openems_microstrip_3d.py:176) and port.CalcPort(sim_path, f); port.uf_ref/port.uf_inc
(antenna_sim/solver_fdtd_openems_microstrip.py:408-413)."""
from __future__ import annotations

import os

import numpy as np

from b200fdtd import postproc


def _ny(d):
    if isinstance(d, str):
        return {"x": 0, "y": 1, "z": 2}[d.lower()]
    d = int(d)
    if d not in (0, 1, 2):
        raise ValueError("port direction must be 0..2 or 'x'|'y'|'z'")
    return d


class Port:
    def __init__(self, CSX, port_nr, start, stop, excite, **kw):
        self.CSX = CSX
        self.number = int(port_nr)
        self.excite = float(excite)
        self.start = np.asarray(start, dtype=np.float64).reshape(3)
        self.stop = np.asarray(stop, dtype=np.float64).reshape(3)
        self.Z_ref = None
        self.U_filenames = kw.get("U_filenames", [])
        self.I_filenames = kw.get("I_filenames", [])
        self.priority = kw.get("priority", 0)
        self.prefix = kw.get("PortNamePrefix", "")
        self.delay = kw.get("delay", 0.0)
        self.lbl_temp = self.prefix + "port_{}" + "_{}".format(self.number)
        self._fdtd = None            # set by openEMS.AddLumpedPort: results live on the FDTD object

    # results of the last Run for this port's probes
    def _records(self, sim_path):
        from . import _registry as _o
        res = _o.results_for(sim_path, self._fdtd)
        if res is None:
            raise RuntimeError(f"CalcPort: no finished simulation found for '{sim_path}'")
        return res

    def ReadUIData(self, sim_path, freq, signal_type="pulse"):
        res = self._records(sim_path)
        pf = res.get("probe_freqs")
        self.uf_tot = 0; self.if_tot = 0
        self.u_data, self.i_data = [], []
        for fn in self.U_filenames:
            pr = res["probes"][fn]
            self.u_data.append(pr); self.uf_tot = self.uf_tot + postproc.port_spectrum(pr, freq, pf)
        for fn in self.I_filenames:
            pr = res["probes"][fn]
            self.i_data.append(pr); self.if_tot = self.if_tot + postproc.port_spectrum(pr, freq, pf)
        self.ut_tot = sum(p["val"] for p in self.u_data)
        self.it_tot = sum(p["val"] for p in self.i_data)
        self.t_u = self.u_data[0]["t"] if self.u_data else None
        self.t_i = self.i_data[0]["t"] if self.i_data else None

    def CalcPort(self, sim_path, freq, ref_impedance=None, ref_plane_shift=None, signal_type="pulse"):
        freq = np.atleast_1d(np.asarray(freq, np.float64))
        self.ReadUIData(sim_path, freq, signal_type)
        if ref_impedance is not None:
            self.Z_ref = ref_impedance
        if self.Z_ref is None:
            raise RuntimeError("Port Z_ref should not be None!")
        if ref_plane_shift is not None:
            raise NotImplementedError("ref_plane_shift needs a transmission-line port (MSL), not used by the reference")
        self.freq = freq
        self.uf_inc = 0.5 * (self.uf_tot + self.if_tot * self.Z_ref)
        self.if_inc = 0.5 * (self.if_tot + self.uf_tot / self.Z_ref)
        self.uf_ref = self.uf_tot - self.uf_inc
        self.if_ref = self.if_inc - self.if_tot
        self.P_inc = 0.5 * np.real(self.uf_inc * np.conj(self.if_inc))
        self.P_ref = 0.5 * np.real(self.uf_ref * np.conj(self.if_ref))
        self.P_acc = 0.5 * np.real(self.uf_tot * np.conj(self.if_tot))


class LumpedPort(Port):
    """Lumped resistor + soft E excitation + voltage/current probes over one box (App. A5)."""

    def __init__(self, CSX, port_nr, R, start, stop, exc_dir, excite=0, **kw):
        super().__init__(CSX, port_nr=port_nr, start=start, stop=stop, excite=excite, **kw)
        self.R = float(R)
        self.exc_ny = _ny(exc_dir)
        self.direction = np.sign(self.stop[self.exc_ny] - self.start[self.exc_ny])
        if not self.start[self.exc_ny] != self.stop[self.exc_ny]:
            raise Exception("LumpedPort: start and stop may not be identical in excitation direction")
        if self.R > 0:
            lumped = CSX.AddLumpedElement(self.lbl_temp.format("resist"), ny=self.exc_ny, caps=True, R=self.R)
        elif self.R == 0:
            lumped = CSX.AddMetal(self.lbl_temp.format("resist"))
        else:
            lumped = None
        if lumped is not None:
            lumped.AddBox(self.start, self.stop, priority=self.priority)
        if self.excite != 0:
            vec = np.zeros(3)
            vec[self.exc_ny] = -1 * self.direction * self.excite
            exc = CSX.AddExcitation(self.lbl_temp.format("excite"), exc_type=0, exc_val=vec, delay=self.delay)
            exc.AddBox(self.start, self.stop, priority=self.priority)
        self.U_filenames = [self.lbl_temp.format("ut")]
        u_start = 0.5 * (self.start + self.stop); u_start[self.exc_ny] = self.start[self.exc_ny]
        u_stop = 0.5 * (self.start + self.stop); u_stop[self.exc_ny] = self.stop[self.exc_ny]
        u_probe = CSX.AddProbe(self.U_filenames[0], p_type=0, weight=-1)
        u_probe.AddBox(u_start, u_stop)
        self.I_filenames = [self.lbl_temp.format("it")]
        mid = 0.5 * (self.start[self.exc_ny] + self.stop[self.exc_ny])
        i_start = np.array(self.start); i_start[self.exc_ny] = mid
        i_stop = np.array(self.stop); i_stop[self.exc_ny] = mid
        i_probe = CSX.AddProbe(self.I_filenames[0], p_type=1, weight=self.direction, norm_dir=self.exc_ny)
        i_probe.AddBox(i_start, i_stop)

    def CalcPort(self, sim_path, freq, ref_impedance=None, ref_plane_shift=None, signal_type="pulse"):
        if ref_impedance is None:
            self.Z_ref = self.R
        super().CalcPort(sim_path, freq, ref_impedance, ref_plane_shift, signal_type)


def write_probe_files(sim_path, results):
    """openEMS-style ASCII probe files `port_ut_1`, `port_it_1` (time, value) under sim_path"""
    for name, pr in results.get("probes", {}).items():
        with open(os.path.join(sim_path, name), "w") as f:
            f.write("% time-domain " + ("voltage" if pr["kind"] == 0 else "current") + " integration by the b200 FDTD engine\n")
            f.write("% t/s\t" + ("voltage" if pr["kind"] == 0 else "current") + "\n")
            for t, v in zip(pr["t"], pr["val"]):
                f.write(f"{t:.12e}\t{v:.9e}\n")
