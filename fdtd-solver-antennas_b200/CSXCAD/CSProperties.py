"""Properties (materials, metals, lumped elements, excitations, probes, dumps) and their primitives.

Call surface used by the reference:
  CSX.AddMaterial('substrate', epsilon=…, kappa=…)     antenna_sim/solver_fdtd_openems_microstrip_3d.py:112
  mat.SetMaterialProperty('Eps'|'Kappa', v)             antenna_sim/solver_fdtd_openems.py:207-211
  CSX.AddMetal('ground'); prop.AddBox(priority=10, start=…, stop=…)            …microstrip_3d.py:118-120
"""
from __future__ import annotations

from .CSPrimitives import CSPrimBox

_MAT_KEYS = {"epsilon": "epsilon", "eps": "epsilon", "mue": "mue", "kappa": "kappa", "sigma": "sigma",
             "density": "density"}


class CSProperties:
    type_name = "Unknown"

    def __init__(self, csx, name="", **kw):
        self.csx = csx
        self.name = name
        self.primitives = []
        self.attrs = dict(kw)

    def GetName(self):
        return self.name

    def GetTypeString(self):
        return self.type_name

    def GetQtyPrimitives(self):
        return len(self.primitives)

    def GetAllPrimitives(self):
        return list(self.primitives)

    def AddBox(self, start=None, stop=None, priority=0, **kw):
        if start is None or stop is None:
            raise ValueError("AddBox needs start and stop")
        prim = CSPrimBox(self, start, stop, priority=priority)
        self.primitives.append(prim)
        self.csx._order += 1
        prim.order = self.csx._order
        return prim

    def _unsupported(self, what):
        raise NotImplementedError(f"{what}: only box primitives are supported by the B200 FDTD shim")

    def AddCylinder(self, *a, **k):
        self._unsupported("AddCylinder")

    def AddPolygon(self, *a, **k):
        self._unsupported("AddPolygon")

    def AddLinPoly(self, *a, **k):
        self._unsupported("AddLinPoly")

    def AddCurve(self, *a, **k):
        self._unsupported("AddCurve")

    def AddWire(self, *a, **k):
        self._unsupported("AddWire")

    def AddSphere(self, *a, **k):
        self._unsupported("AddSphere")


class CSPropMaterial(CSProperties):
    type_name = "Material"

    def __init__(self, csx, name="", **kw):
        super().__init__(csx, name)
        self.props = {"epsilon": 1.0, "mue": 1.0, "kappa": 0.0, "sigma": 0.0, "density": 0.0}
        self.SetMaterialProperty(**kw)

    def SetMaterialProperty(self, *args, **kw):
        # both spellings occur: SetMaterialProperty(epsilon=4.3) and SetMaterialProperty('Eps', 4.3)
        if len(args) == 2 and isinstance(args[0], str):
            kw = dict(kw); kw[args[0]] = args[1]
        elif args:
            raise TypeError("SetMaterialProperty(name, value) or SetMaterialProperty(key=value)")
        for k, v in kw.items():
            key = _MAT_KEYS.get(str(k).lower())
            if key is None:
                raise ValueError(f"unknown material property '{k}'")
            self.props[key] = float(v)
        if self.props["mue"] != 1.0 or self.props["sigma"] != 0.0:
            raise NotImplementedError("magnetic materials (mue != 1, sigma != 0) are outside the reference's path")

    def GetMaterialProperty(self, name):
        return self.props[_MAT_KEYS[str(name).lower()]]


class CSPropMetal(CSProperties):
    type_name = "Metal"


class CSPropLumpedElement(CSProperties):
    type_name = "LumpedElement"

    def __init__(self, csx, name="", ny=2, caps=True, R=None, C=None, L=None, **kw):
        super().__init__(csx, name)
        self.ny = _dir(ny)
        self.caps = bool(caps)
        self.R = R
        self.C = C
        self.L = L
        if L not in (None, 0) or C not in (None, 0):
            raise NotImplementedError("lumped L/C elements are outside the reference's path (only R is used)")


class CSPropExcitation(CSProperties):
    type_name = "Excitation"

    def __init__(self, csx, name="", exc_type=0, exc_val=(0, 0, 0), delay=0.0, **kw):
        super().__init__(csx, name)
        self.exc_type = int(exc_type)
        self.exc_val = [float(v) for v in exc_val]
        self.delay = float(delay)
        if self.exc_type != 0:
            raise NotImplementedError("only soft E-field excitation (exc_type=0) is used by the reference's ports")


class CSPropProbeBox(CSProperties):
    type_name = "ProbeBox"

    def __init__(self, csx, name="", p_type=0, weight=1.0, norm_dir=-1, **kw):
        super().__init__(csx, name)
        self.p_type = int(p_type)
        self.weight = float(weight)
        self.norm_dir = int(norm_dir)


class CSPropDumpBox(CSProperties):
    type_name = "DumpBox"

    def __init__(self, csx, name="", dump_type=0, dump_mode=1, file_type=1, frequency=None, **kw):
        super().__init__(csx, name)
        self.dump_type = int(dump_type)
        self.dump_mode = int(dump_mode)
        self.file_type = int(file_type)
        self.frequency = None if frequency is None else [float(f) for f in frequency]


def _dir(d):
    if isinstance(d, str):
        return {"x": 0, "y": 1, "z": 2}[d.lower()]
    d = int(d)
    if d not in (0, 1, 2):
        raise ValueError(f"direction {d} not in 0..2")
    return d
