"""openEMS.automesh stand-in: mesh hints from box primitives (App. A8 thirds rule)."""
import numpy as np

from b200fdtd.mesh import edges_to_lines


def mesh_hint_from_box(box, dirs, metal_edge_res=None, **kw):
    hint = [None, None, None]
    start, stop = np.asarray(box.GetStart(), float), np.asarray(box.GetStop(), float)
    for a in _dirs(dirs):
        hint[a] = edges_to_lines(start[a], stop[a], metal_edge_res)
    return hint


def mesh_hint_from_primitive(primitive, dirs, **kw):
    if primitive.GetTypeName() != "Box":
        return None
    if primitive.HasTransform():
        # transformed primitives are skipped, like openEMS (reference works around it: …multi_3d.py:309-324)
        return None
    return mesh_hint_from_box(primitive, dirs, **kw)


def mesh_combine(mesh1, mesh2, sort=True):
    mesh = [None, None, None]
    for a in range(3):
        if mesh1[a] is None and mesh2[a] is None:
            continue
        l = list(mesh1[a] or []) + list(mesh2[a] or [])
        mesh[a] = sorted(l) if sort else l
    return mesh


def _dirs(dirs):
    if isinstance(dirs, str):
        if dirs == "all":
            return [0, 1, 2]
        return ["xyz".index(c) for c in dirs.lower()]
    return [int(d) for d in np.atleast_1d(dirs)]
