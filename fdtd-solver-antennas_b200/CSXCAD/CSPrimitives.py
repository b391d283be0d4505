"""Geometric primitives.  The reference only creates boxes: prop.AddBox(start, stop, priority=…)
(antenna_sim/solver_fdtd_openems_microstrip_3d.py:115,120,127,156; …multi_3d.py:357-445)."""
from __future__ import annotations

import numpy as np

from .CSTransform import CSTransform


class CSPrimitives:
    def __init__(self, prop, priority=0):
        self.prop = prop
        self.priority = int(priority)
        self.transform = None

    def GetProperty(self):
        return self.prop

    def GetPriority(self):
        return self.priority

    def SetPriority(self, v):
        self.priority = int(v)

    def AddTransform(self, name, *args, **kw):
        if self.transform is None:
            self.transform = CSTransform()
        self.transform.AddTransform(name, *args, **kw)

    def GetTransform(self):
        return self.transform

    def HasTransform(self):
        return self.transform is not None and self.transform.HasTransform()


class CSPrimBox(CSPrimitives):
    def __init__(self, prop, start, stop, priority=0):
        super().__init__(prop, priority)
        self.start = np.asarray(start, dtype=np.float64).reshape(3).copy()
        self.stop = np.asarray(stop, dtype=np.float64).reshape(3).copy()

    def GetType(self):
        return 1  # BOX

    def GetTypeName(self):
        return "Box"

    def GetStart(self):
        return self.start.copy()

    def GetStop(self):
        return self.stop.copy()

    def SetStart(self, v):
        self.start = np.asarray(v, dtype=np.float64).reshape(3).copy()

    def SetStop(self, v):
        self.stop = np.asarray(v, dtype=np.float64).reshape(3).copy()

    def lo(self):
        return np.minimum(self.start, self.stop)

    def hi(self):
        return np.maximum(self.start, self.stop)

    def GetBoundBox(self):
        """axis-aligned bounding box in world coordinates (corners transformed if needed)"""
        lo, hi = self.lo(), self.hi()
        if not self.HasTransform():
            return np.array([lo, hi])
        c = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
        w = c @ self.transform.M.T + self.transform.t
        return np.array([w.min(0), w.max(0)])

    def contains(self, pts, tol=0.0):
        """pts [..., 3] world coordinates -> boolean mask (closed box, like CSXCAD's IsInside)"""
        p = np.asarray(pts, dtype=np.float64)
        if self.HasTransform():
            p = self.transform.to_local(p)
        lo, hi = self.lo() - tol, self.hi() + tol
        return np.all((p >= lo) & (p <= hi), axis=-1)
