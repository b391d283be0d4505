"""Small scenes built through the CSXCAD/openEMS shim API (test helper), runnable on either engine."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fdtd-solver-antennas_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def use_oracle_engine(threads=4):
    """route the shim's Run through the CPU oracle (tests only)"""
    from oracle.fdtd_ref import RefEngine, farfield
    from openEMS import openEMS as O
    O.default_engine_factory = staticmethod(lambda nx, ny, nz, px, dev: RefEngine(nx, ny, nz, px, threads=threads))
    O.default_farfield_fn = staticmethod(farfield)


def use_cuda_engine():
    from openEMS import openEMS as O
    O.default_engine_factory = None
    O.default_farfield_fn = None


def tmp_sim_path(tag):
    return os.path.join(tempfile.gettempdir(), f"b200fdtd_{tag}_{os.getpid()}")


def cavity(a=0.05, b=0.04, c=0.03, n=(21, 17, 13), eps_r=1.0, nrts=4000, f0=5e9, fc=3e9):
    from CSXCAD import ContinuousStructure
    from openEMS import openEMS
    F = openEMS(NrTS=nrts, EndCriteria=1e-12)
    F.SetGaussExcite(f0, fc)
    F.SetBoundaryCond(["PEC"] * 6)
    csx = ContinuousStructure(); F.SetCSX(csx)
    g = csx.GetGrid(); g.SetDeltaUnit(1.0)
    lines = [np.linspace(0, L, m) for L, m in zip((a, b, c), n)]
    for ax, l in enumerate(lines):
        g.AddLine("xyz"[ax], l)
    if eps_r != 1.0:
        csx.AddMaterial("fill", epsilon=eps_r).AddBox([0, 0, 0], [a, b, c])
    x, y, z = lines
    ex = csx.AddExcitation("src", 0, [0, 0, 1])
    ex.AddBox([x[6], y[5], z[3]], [x[6], y[5], z[5]])
    pr = csx.AddProbe("ut_cav", 0)
    pr.AddBox([x[13], y[10], z[6]], [x[13], y[10], z[8]])
    return F


def dipole(boundary="PML_8", cells=(40, 40, 48), delta=2.5e-3, f0=3e9, fc=1.5e9, nrts=2500, end=1e-5, R=50.0,
           arm_edges=4, nf2ff_freqs=None):
    """z-directed centre-fed PEC dipole with a lumped port, in vacuum"""
    from CSXCAD import ContinuousStructure
    from openEMS import openEMS
    F = openEMS(NrTS=nrts, EndCriteria=end)
    F.SetGaussExcite(f0, fc)
    F.SetBoundaryCond([boundary] * 6)
    csx = ContinuousStructure(); F.SetCSX(csx)
    g = csx.GetGrid(); g.SetDeltaUnit(1e-3)
    d = delta * 1e3
    nx, ny, nz = cells
    g.AddLine("x", (np.arange(nx + 1) - nx / 2) * d)
    g.AddLine("y", (np.arange(ny + 1) - ny / 2) * d)
    g.AddLine("z", (np.arange(nz + 1) - nz / 2 + 0.5) * d)          # an edge is centred on z=0
    m = csx.AddMetal("arms")
    m.AddBox([0, 0, 0.5 * d], [0, 0, (0.5 + arm_edges) * d], priority=10)
    m.AddBox([0, 0, -0.5 * d], [0, 0, -(0.5 + arm_edges) * d], priority=10)
    port = F.AddLumpedPort(1, R, [0, 0, -0.5 * d], [0, 0, 0.5 * d], "z", 1.0, priority=5)
    if nf2ff_freqs is not None:
        F.nf2ff_freqs = nf2ff_freqs
    nf = F.CreateNF2FFBox()
    return F, nf, port
