"""Synthetic workloads of the BASELINE.json configs, built through the drop-in CSXCAD/openEMS API.

`patch_scene` is the 2.45 GHz FR-4 microstrip-fed patch of the reference's single-patch backend
(scene recipe: antenna_sim/solver_fdtd_openems_microstrip_3d.py:43-194, feed direction -X) with the mesh
resolution given explicitly, so the same scene can be meshed at ~0.3 M cells (config 1), ~100 M cells
(config 2) or more (weak scaling).  Dimensions are the closed-form values the reference's own helpers produce
for f = 2.45 GHz, eps_r = 4.3, h = 1.6 mm (pinned in tests/golden/closed_form.json).

`vacuum_cube` is config 5: uniform vacuum, Mur on all faces, one z-directed soft source in the centre.
"""
from __future__ import annotations

import math

import numpy as np

from .constants import C0, EPS0

# closed-form design values of config 1 (mm), SURVEY.md App. B
PATCH_W_MM = 37.58388632919335       # along x (feed axis)
PATCH_L_MM = 29.138326192715315      # along y
FEED_W_MM = 3.1143958916679416
SUB_H_MM = 1.6
EPS_R = 4.3
TAN_D = 0.02
F_DESIGN = 2.45e9


def patch_scene(mesh_res_mm=None, target_cells=None, boundary="PML_8", f0=2.45e9, fc=1.225e9, nrts=30000,
                end_criteria=1e-4, nf2ff_freqs=None, z_refine=1):
    """returns (FDTD, nf2ff, port).  Either mesh_res_mm or target_cells fixes the resolution."""
    from CSXCAD import ContinuousStructure
    from openEMS import openEMS

    unit = 1e-3
    h = SUB_H_MM
    margin, feed_len, air = 30.0, 20.0, 80.0
    sub_x = PATCH_W_MM + 2 * margin + feed_len
    sub_y = PATCH_L_MM + 2 * margin
    box = (sub_x + 2 * air, sub_y + 2 * air, 160.0)
    if mesh_res_mm is None:
        if target_cells is None:
            mesh_res_mm = C0 / (f0 + fc) / unit / 20.0               # quality 3 of the reference
        else:
            mesh_res_mm = (box[0] * box[1] * box[2] / float(target_cells)) ** (1.0 / 3.0)
    res = float(mesh_res_mm)

    FDTD = openEMS(NrTS=nrts, EndCriteria=end_criteria)
    FDTD.SetGaussExcite(f0, fc)
    FDTD.SetBoundaryCond([boundary] * 6)
    CSX = ContinuousStructure()
    FDTD.SetCSX(CSX)
    mesh = CSX.GetGrid()
    mesh.SetDeltaUnit(unit)
    mesh.AddLine("x", [-box[0] / 2, box[0] / 2])
    mesh.AddLine("y", [-box[1] / 2, box[1] / 2])
    mesh.AddLine("z", [-box[2] / 3, box[2] * 2 / 3])

    kappa = 2 * math.pi * F_DESIGN * EPS0 * EPS_R * TAN_D
    sub = CSX.AddMaterial("substrate", epsilon=EPS_R, kappa=kappa)
    s0, s1 = [-sub_x / 2, -sub_y / 2, 0.0], [sub_x / 2, sub_y / 2, h]
    sub.AddBox(priority=0, start=s0, stop=s1)
    n_sub = max(4, int(round(h / res)) * max(1, int(z_refine)))
    mesh.AddLine("z", np.linspace(0, h, n_sub + 1))

    gnd = CSX.AddMetal("ground")
    gnd.AddBox(priority=10, start=[s0[0], s0[1], 0.0], stop=[s1[0], s1[1], 0.0])
    FDTD.AddEdges2Grid(dirs="xy", properties=gnd)
    patch = CSX.AddMetal("patch")
    patch.AddBox(priority=10, start=[-PATCH_W_MM / 2, -PATCH_L_MM / 2, h], stop=[PATCH_W_MM / 2, PATCH_L_MM / 2, h])
    FDTD.AddEdges2Grid(dirs="xy", properties=patch, metal_edge_res=res / 2)
    feed = CSX.AddMetal("feed_line")
    feed.AddBox(priority=10, start=[-sub_x / 2, -FEED_W_MM / 2, h], stop=[-PATCH_W_MM / 2, FEED_W_MM / 2, h])
    FDTD.AddEdges2Grid(dirs="xy", properties=feed, metal_edge_res=res / 2)

    fx, fy = -PATCH_W_MM / 2, 0.0
    mesh.AddLine("x", [fx]); mesh.AddLine("y", [fy]); mesh.AddLine("z", [0.0, h])
    port = FDTD.AddLumpedPort(1, 50.0, [fx, fy, 0.0], [fx, fy, h], "z", 1.0, priority=5, edges2grid="xy")
    mesh.SmoothMeshLines("all", res, 1.4)
    if nf2ff_freqs is not None:
        FDTD.nf2ff_freqs = nf2ff_freqs
    nf2ff = FDTD.CreateNF2FFBox()
    return FDTD, nf2ff, port


def vacuum_cube(n, delta=1e-3, boundary="MUR", f0=10e9, fc=5e9, nrts=1000, nz=None):
    """config 5: n x n x nz uniform vacuum grid, z-directed soft source on the centre edge"""
    from CSXCAD import ContinuousStructure
    from openEMS import openEMS
    nz = int(nz or n)
    FDTD = openEMS(NrTS=nrts, EndCriteria=1e-12)
    FDTD.SetGaussExcite(f0, fc)
    FDTD.SetBoundaryCond([boundary] * 6)
    CSX = ContinuousStructure()
    FDTD.SetCSX(CSX)
    g = CSX.GetGrid()
    g.SetDeltaUnit(1.0)
    g.AddLine("x", np.arange(n) * delta)
    g.AddLine("y", np.arange(n) * delta)
    g.AddLine("z", np.arange(nz) * delta)
    c = (n // 2) * delta
    cz = (nz // 2) * delta
    ex = CSX.AddExcitation("src", 0, [0, 0, 1])
    ex.AddBox([c, c, cz], [c, c, cz + delta])
    pr = CSX.AddProbe("ut_centre", 0, weight=-1)
    pr.AddBox([c, c + 8 * delta, cz], [c, c + 8 * delta, cz + delta])
    return FDTD
