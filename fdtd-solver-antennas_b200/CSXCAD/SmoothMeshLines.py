"""CSXCAD.SmoothMeshLines module stand-in."""
from b200fdtd.mesh import smooth_mesh_lines as _s


def SmoothMeshLines(lines, max_res, ratio=1.5, **kw):
    return _s(lines, max_res, ratio)
