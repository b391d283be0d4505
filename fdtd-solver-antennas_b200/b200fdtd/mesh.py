"""Rectilinear mesh helpers: line smoothing and metal-edge refinement.

Behavioural contract of the CSXCAD/openEMS helpers the reference calls:
  mesh.SmoothMeshLines('all', mesh_res, 1.4)           antenna_sim/solver_fdtd_openems_microstrip_3d.py:178
  FDTD.AddEdges2Grid(dirs='xy', properties=p, metal_edge_res=r)      …microstrip_3d.py:122,128,157
Lines are an INPUT of the engine: parity runs feed the oracle and the CUDA engine the same lines.
"""
from __future__ import annotations

import numpy as np


def unique_lines(lines, tol=1e-9):
    """sorted lines with (near-)duplicates removed (reference passes repeated values: …multi_3d.py:366-378)"""
    a = np.sort(np.asarray(lines, dtype=np.float64).ravel())
    if a.size == 0:
        return a
    scale = max(1.0, float(np.abs(a).max()))
    keep = np.concatenate([[True], np.diff(a) > tol * scale])
    return a[keep]


def _fill_gap(a, b, d_left, d_right, max_res, ratio):
    """interior points for the gap (a,b) whose neighbours have steps d_left / d_right (None at the mesh end)"""
    L = b - a
    if L <= max_res * (1 + 1e-9):
        return []
    sl = min(max_res, d_left * ratio) if d_left else max_res
    sr = min(max_res, d_right * ratio) if d_right else max_res
    left, right = [], []
    total = 0.0
    while total < L:
        if sl <= sr:
            left.append(sl); total += sl; sl = min(max_res, sl * ratio)
        else:
            right.append(sr); total += sr; sr = min(max_res, sr * ratio)
    steps = np.array(left + right[::-1])
    if len(steps) > 1 and total - L > 0.5 * steps.min():
        # dropping the smallest step and stretching would break max_res; shrink everything instead
        pass
    steps *= L / total
    pts = a + np.cumsum(steps)[:-1]
    return list(pts)


def smooth_mesh_lines(lines, max_res, ratio=1.4):
    """Insert lines so that every step is <= max_res and neighbouring steps differ by <= ratio
    (graded from the existing fine regions).  Existing lines are kept."""
    out = list(unique_lines(lines))
    if len(out) < 2:
        return np.array(out)
    guard = 0
    while guard < 100000:
        guard += 1
        d = np.diff(out)
        big = np.where(d > max_res * (1 + 1e-9))[0]
        if big.size == 0:
            break
        # work on the gap whose neighbours are finest first (so grading propagates outwards)
        def nb(i):
            l = d[i - 1] if i > 0 else np.inf
            r = d[i + 1] if i + 1 < len(d) else np.inf
            return min(l, r)
        i = min(big, key=nb)
        dl = d[i - 1] if i > 0 and d[i - 1] <= max_res * (1 + 1e-9) else None
        dr = d[i + 1] if i + 1 < len(d) and d[i + 1] <= max_res * (1 + 1e-9) else None
        pts = _fill_gap(out[i], out[i + 1], dl, dr, max_res, ratio)
        out[i + 1:i + 1] = pts
    return np.array(out)


def edges_to_lines(start, stop, metal_edge_res=None):
    """lines contributed by one box edge pair along one axis (thirds rule when metal_edge_res is given)"""
    lo, hi = min(start, stop), max(start, stop)
    if metal_edge_res is None or metal_edge_res <= 0 or hi == lo:
        return [lo, hi] if hi != lo else [lo]
    m = float(metal_edge_res)
    return [lo - 2.0 * m / 3.0, lo + m / 3.0, hi - m / 3.0, hi + 2.0 * m / 3.0]
