"""ContinuousStructure: the scene container (antenna_sim/solver_fdtd_openems_microstrip_3d.py:92-93)."""
from __future__ import annotations

from . import CSProperties as _P
from .CSRectGrid import CSRectGrid


class ContinuousStructure:
    def __init__(self, CoordSystem=0, **kw):
        if int(CoordSystem) != 0:
            raise NotImplementedError("only Cartesian coordinates are supported")
        self.grid = CSRectGrid()
        self.properties = []
        self._order = 0

    def GetGrid(self):
        return self.grid

    def _add(self, prop):
        self.properties.append(prop)
        return prop

    def AddMaterial(self, name, **kw):
        return self._add(_P.CSPropMaterial(self, name, **kw))

    def AddMetal(self, name, **kw):
        return self._add(_P.CSPropMetal(self, name))

    def AddLumpedElement(self, name, **kw):
        return self._add(_P.CSPropLumpedElement(self, name, **kw))

    def AddExcitation(self, name, exc_type, exc_val, **kw):
        return self._add(_P.CSPropExcitation(self, name, exc_type=exc_type, exc_val=exc_val, **kw))

    def AddProbe(self, name, p_type, **kw):
        return self._add(_P.CSPropProbeBox(self, name, p_type=p_type, **kw))

    def AddDump(self, name, **kw):
        return self._add(_P.CSPropDumpBox(self, name, **kw))

    def GetAllProperties(self):
        return list(self.properties)

    def GetQtyProperties(self):
        return len(self.properties)

    def GetPropertiesByType(self, cls):
        return [p for p in self.properties if isinstance(p, cls)]

    def GetPropertiesByName(self, name):
        return [p for p in self.properties if p.name == name]

    def GetAllPrimitives(self, sort=False, prop_type=None):
        prims = [q for p in self.properties if prop_type is None or isinstance(p, prop_type) for q in p.primitives]
        if sort:
            prims.sort(key=lambda q: (-q.priority, q.order))
        return prims

    def GetQtyPrimitives(self):
        return sum(len(p.primitives) for p in self.properties)

    def Write2XML(self, fn):
        """debug artefact only: a plain-text listing (CSXCAD's XML schema is not reproduced)"""
        with open(fn, "w") as f:
            f.write("<!-- b200 FDTD shim scene listing -->\n")
            for a, n in enumerate("xyz"):
                f.write(f"<{n}lines unit='{self.grid.unit}'>{','.join(repr(float(v)) for v in self.grid.lines[a])}</{n}lines>\n")
            for p in self.properties:
                for q in p.primitives:
                    f.write(f"<prim type='{p.type_name}' name='{p.name}' prio='{q.priority}' "
                            f"start='{q.start.tolist()}' stop='{q.stop.tolist()}'/>\n")
