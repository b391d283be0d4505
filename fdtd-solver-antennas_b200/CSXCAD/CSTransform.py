"""Primitive transforms: prim.AddTransform('RotateAxis', 'x'|'y'|'z', deg) and
prim.AddTransform('Translate', [x, y, z])  (antenna_sim/solver_fdtd_openems_microstrip_multi_3d.py:359-363)."""
from __future__ import annotations

import numpy as np

_AX = {"x": 0, "y": 1, "z": 2, 0: 0, 1: 1, 2: 2}


class CSTransform:
    """Affine transform p_world = M @ p_local + t, composed in the order transforms are added."""

    def __init__(self):
        self.M = np.eye(3)
        self.t = np.zeros(3)
        self.ops = []

    def HasTransform(self):
        return len(self.ops) > 0

    def AddTransform(self, name, *args, **kw):
        name = str(name)
        if name == "Translate":
            v = np.asarray(args[0], dtype=np.float64).reshape(3)
            self.t = self.t + v
            self.ops.append(("Translate", v.tolist()))
        elif name == "RotateAxis":
            ax = _AX[args[0].lower() if isinstance(args[0], str) else int(args[0])]
            ang = float(args[1])
            if kw.get("deg", True):
                ang = np.deg2rad(ang)
            c, s = np.cos(ang), np.sin(ang)
            R = np.eye(3)
            a, b = (ax + 1) % 3, (ax + 2) % 3
            R[a, a] = c; R[a, b] = -s; R[b, a] = s; R[b, b] = c
            self.M = R @ self.M
            self.t = R @ self.t
            self.ops.append(("RotateAxis", "xyz"[ax], float(args[1])))
        elif name == "Scale":
            f = np.asarray(args[0], dtype=np.float64)
            S = np.diag(np.broadcast_to(f, (3,)).astype(np.float64))
            self.M = S @ self.M
            self.t = S @ self.t
            self.ops.append(("Scale", np.broadcast_to(f, (3,)).tolist()))
        else:
            raise ValueError(f"CSTransform: unsupported transform '{name}'")

    def Transform(self, p):
        return self.M @ np.asarray(p, dtype=np.float64) + self.t

    def to_local(self, pts):
        """pts [..., 3] world -> local coordinates"""
        Minv = np.linalg.inv(self.M)
        return (np.asarray(pts, dtype=np.float64) - self.t) @ Minv.T
