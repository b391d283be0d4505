"""Rectilinear grid (CSX.GetGrid()): SetDeltaUnit / AddLine / SmoothMeshLines
(antenna_sim/solver_fdtd_openems_microstrip_3d.py:94-95,107-109,116,171-173,178)."""
from __future__ import annotations

import numpy as np

from b200fdtd import mesh as _mesh

_AX = {"x": 0, "y": 1, "z": 2, 0: 0, 1: 1, 2: 2}


def _axes(ny):
    if isinstance(ny, str):
        s = ny.lower()
        if s == "all":
            return [0, 1, 2]
        return [_AX[c] for c in s]
    return [_AX[int(ny)]]


class CSRectGrid:
    def __init__(self):
        self.unit = 1.0
        self.lines = [np.zeros(0), np.zeros(0), np.zeros(0)]

    def SetDeltaUnit(self, unit):
        self.unit = float(unit)

    def GetDeltaUnit(self):
        return self.unit

    def SetMeshType(self, t):
        if int(t) != 0:
            raise NotImplementedError("only Cartesian meshes are supported")

    def GetMeshType(self):
        return 0

    def Clear(self):
        self.lines = [np.zeros(0), np.zeros(0), np.zeros(0)]

    def ClearLines(self, ny):
        for a in _axes(ny):
            self.lines[a] = np.zeros(0)

    def AddLine(self, ny, line):
        vals = np.atleast_1d(np.asarray(line, dtype=np.float64)).ravel()
        if not np.all(np.isfinite(vals)):
            raise ValueError("mesh lines must be finite")
        for a in _axes(ny):
            self.lines[a] = np.concatenate([self.lines[a], vals])

    def SetLines(self, ny, lines):
        for a in _axes(ny):
            self.lines[a] = np.atleast_1d(np.asarray(lines, dtype=np.float64)).ravel().copy()

    def Sort(self, ny="all"):
        for a in _axes(ny):
            self.lines[a] = _mesh.unique_lines(self.lines[a])

    def GetLines(self, ny, do_sort=False):
        a = _axes(ny)[0]
        if do_sort:
            self.Sort(a)
        return self.lines[a].copy()

    def GetLine(self, ny, idx):
        return float(self.lines[_axes(ny)[0]][idx])

    def GetQtyLines(self, ny):
        return int(len(self.lines[_axes(ny)[0]]))

    def SmoothMeshLines(self, ny, max_res, ratio=1.5):
        for a in _axes(ny):
            self.lines[a] = _mesh.smooth_mesh_lines(self.lines[a], float(max_res), float(ratio))

    def GetSimArea(self):
        self.Sort()
        return np.array([[l[0] for l in self.lines], [l[-1] for l in self.lines]])

    def IsValid(self):
        self.Sort()
        return all(len(l) >= 2 for l in self.lines)
