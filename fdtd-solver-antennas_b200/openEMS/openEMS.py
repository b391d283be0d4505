"""openEMS.openEMS — the FDTD front object of the reference's solver backends, driving the B200 engine.

Call surface (antenna_sim/solver_fdtd_openems_microstrip_3d.py:82-93,122,128,157,176,179,214):
  openEMS(NrTS=…, EndCriteria=…), SetGaussExcite, SetBoundaryCond, SetCSX, AddEdges2Grid,
  AddLumpedPort, CreateNF2FFBox, Run(sim_path, verbose=, cleanup=)
"""
from __future__ import annotations

import os
import shutil
import sys

import numpy as np

from b200fdtd import mesh as _mesh
from b200fdtd.operator import Setup, parse_bc, BC_MUR, BC_PML, BC_PEC
from b200fdtd.simulation import Simulation
from CSXCAD import CSProperties as _P
from . import ports as _ports
from . import nf2ff as _nf2ff

from . import _registry


class _PrimAdapter:
    """primitive in SI units for the operator builder"""

    def __init__(self, prim, unit):
        self.prim = prim
        self.unit = unit
        self.axis_aligned = not prim.HasTransform()
        self.bbox_m = prim.GetBoundBox() * unit

    def contains_m(self, pts_m, tol):
        return self.prim.contains(np.asarray(pts_m) / self.unit, tol / self.unit)


class openEMS:
    # test hooks: the CPU suite injects the oracle engine / far-field here; the product default (None) is the
    # CUDA engine, which raises if no GPU or no built library is present (no CPU fallback)
    default_engine_factory = None
    default_farfield_fn = None

    def __init__(self, NrTS=1e9, EndCriteria=1e-5, MaxTime=0, OverSampling=4, CoordSystem=0, MultiGrid=None,
                 TimeStepMethod=None, TimeStepFactor=1.0, CellConstantMaterial=False, **kw):
        if int(CoordSystem) != 0:
            raise NotImplementedError("only Cartesian FDTD is supported")
        self.NrTS = int(NrTS)
        self.EndCriteria = float(EndCriteria)
        self.OverSampling = int(OverSampling)
        self.TimeStepFactor = float(TimeStepFactor)
        self.CSX = None
        self.bc = None
        self.exc = None
        self.ports = []
        self.nf2ff_boxes = []
        self.results = None
        self.sim = None
        self._prepared = None         # (scene key, Simulation) left by Run(setup_only=True)
        # engine options of the shim (not part of openEMS): GPU index, frequency lists for the running DFTs
        self.device = int(os.environ.get("B200FDTD_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        self.port_dft_freqs = None
        self.nf2ff_freqs = None
        self.nf2ff_td = None          # None: keep the NF2FF face samples in HBM if they fit (far field at any frequency); False: never
        self.engine_factory = type(self).default_engine_factory
        self.farfield_fn = type(self).default_farfield_fn

    # ---- configuration ----
    def SetNumberOfTimeSteps(self, n):
        self.NrTS = int(n)

    def SetEndCriteria(self, v):
        self.EndCriteria = float(v)

    def SetOverSampling(self, v):
        self.OverSampling = int(v)

    def SetTimeStepFactor(self, v):
        self.TimeStepFactor = float(v)

    def SetMaxTime(self, v):
        pass

    def SetGaussExcite(self, f0, fc):
        if fc <= 0:
            raise ValueError("SetGaussExcite: fc must be > 0")
        self.exc = ("gauss", float(f0), float(fc))

    def SetBoundaryCond(self, BC):
        self.bc = parse_bc(list(BC))

    def SetCSX(self, CSX):
        self.CSX = CSX

    def GetCSX(self):
        return self.CSX

    def AddEdges2Grid(self, dirs, primitives=None, properties=None, **kw):
        """add the box edges of the given properties/primitives as mesh lines (thirds rule with metal_edge_res)"""
        if self.CSX is None:
            raise RuntimeError("AddEdges2Grid: CSX is not set")
        prims = []
        if primitives is not None:
            prims += list(primitives) if isinstance(primitives, (list, tuple)) else [primitives]
        if properties is not None:
            props = list(properties) if isinstance(properties, (list, tuple)) else [properties]
            for p in props:
                prims += p.GetAllPrimitives()
        if primitives is None and properties is None:
            prims = self.CSX.GetAllPrimitives()
        grid = self.CSX.GetGrid()
        axes = [0, 1, 2] if dirs == "all" else ["xyz".index(c) for c in str(dirs).lower()]
        mer = kw.get("metal_edge_res", None)
        for prim in prims:
            if prim.HasTransform():
                continue
            use_mer = mer if isinstance(prim.GetProperty(), _P.CSPropMetal) else None
            for a in axes:
                grid.AddLine(a, _mesh.edges_to_lines(prim.start[a], prim.stop[a], use_mer))

    def AddLumpedPort(self, port_nr, R, start, stop, p_dir, excite=0, edges2grid=None, **kw):
        if self.CSX is None:
            raise RuntimeError("AddLumpedPort: CSX is not set")
        port = _ports.LumpedPort(self.CSX, port_nr, R, start, stop, p_dir, excite, **kw)
        port._fdtd = self
        if edges2grid is not None:
            grid = self.CSX.GetGrid()
            for c in str(edges2grid).lower():
                a = "xyz".index(c)
                grid.AddLine(a, np.unique([port.start[a], port.stop[a]]))
        self.ports.append(port)
        return port

    def CreateNF2FFBox(self, name="nf2ff", start=None, stop=None, **kw):
        """Huygens box 2 lines (MUR) / pml+1 lines (PML) inside each boundary unless start/stop are given (App. A6)"""
        if self.CSX is None:
            raise RuntimeError("CreateNF2FFBox: CSX is not set")
        if start is None or stop is None:
            if self.bc is None:
                raise RuntimeError("CreateNF2FFBox: set the boundary conditions first")
            grid = self.CSX.GetGrid()
            types, cells = self.bc
            start, stop = np.zeros(3), np.zeros(3)
            for a in range(3):
                l = grid.GetLines(a, do_sort=True)
                def inset(t, c):
                    return 2 if t == BC_MUR else (c + 1 if t == BC_PML else 0)
                lo, hi = inset(types[2 * a], cells[2 * a]), inset(types[2 * a + 1], cells[2 * a + 1])
                if lo + hi + 2 > len(l):
                    raise RuntimeError("CreateNF2FFBox: mesh too small for the boundary inset")
                start[a] = l[lo]; stop[a] = l[-hi - 1]
        box = _nf2ff.nf2ff(self.CSX, name, start, stop, **kw)
        box._fdtd = self
        self.nf2ff_boxes.append(box)
        return box

    # ---- scene -> Setup ----
    def _setup(self):
        if self.CSX is None:
            raise RuntimeError("Run: no CSX set (SetCSX)")
        if self.bc is None:
            raise RuntimeError("Run: no boundary conditions set (SetBoundaryCond)")
        if self.exc is None:
            raise RuntimeError("Run: no excitation set (SetGaussExcite)")
        grid = self.CSX.GetGrid()
        unit = grid.GetDeltaUnit()
        lines = [_mesh.unique_lines(grid.GetLines(a)) * unit for a in range(3)]
        types, cells = self.bc
        S = Setup(lines=lines, bc=types, pml_cells=cells, f0=self.exc[1], fc=self.exc[2], nrts=self.NrTS,
                  end_criteria=self.EndCriteria, timestep_factor=self.TimeStepFactor, oversampling=self.OverSampling)
        for prop in self.CSX.GetAllProperties():
            for prim in prop.GetAllPrimitives():
                ad = _PrimAdapter(prim, unit)
                lo, hi = ad.bbox_m[0], ad.bbox_m[1]
                if isinstance(prop, _P.CSPropMaterial):
                    S.materials.append(dict(eps=prop.props["epsilon"], kappa=prop.props["kappa"], priority=prim.priority,
                                            order=prim.order, prim=ad))
                elif isinstance(prop, _P.CSPropMetal):
                    S.metals.append(dict(priority=prim.priority, order=prim.order, prim=ad))
                elif isinstance(prop, _P.CSPropLumpedElement):
                    if not ad.axis_aligned:
                        raise NotImplementedError("transformed lumped elements are not supported")
                    S.lumped.append(dict(ny=prop.ny, R=prop.R, caps=prop.caps, lo=lo, hi=hi))
                elif isinstance(prop, _P.CSPropExcitation):
                    S.excitations.append(dict(vec=prop.exc_val, delay=prop.delay, lo=lo, hi=hi))
                elif isinstance(prop, _P.CSPropProbeBox):
                    S.probes.append(dict(name=prop.name, p_type=prop.p_type, weight=prop.weight, norm_dir=prop.norm_dir,
                                         start=prim.start * unit, stop=prim.stop * unit))
        if self.nf2ff_boxes:
            b = self.nf2ff_boxes[0]
            S.nf2ff = dict(start=b.start * unit, stop=b.stop * unit, frequency=b.freq)
        f0, fc = self.exc[1], self.exc[2]
        S.probe_freqs = (np.asarray(self.port_dft_freqs, np.float64) if self.port_dft_freqs is not None
                         else np.linspace(max(f0 - fc, 0.0), f0 + fc, 201))
        return S

    def _scene_key(self, S, world):
        """fingerprint of everything the operator and the run length depend on (a prepared set-up is reused only for this scene)"""
        import hashlib
        h = hashlib.sha1()
        for l in S.lines:
            h.update(np.ascontiguousarray(l, np.float64).tobytes())
        def box(p):
            t = p.prim.GetTransform()
            tr = b"" if t is None else np.ascontiguousarray(t.M, np.float64).tobytes() + np.ascontiguousarray(t.t, np.float64).tobytes()
            return (np.ascontiguousarray(p.prim.start, np.float64).tobytes() + np.ascontiguousarray(p.prim.stop, np.float64).tobytes() + tr)
        for m in S.materials:
            h.update(repr((m["eps"], m["kappa"], m["priority"], m["order"])).encode() + box(m["prim"]))
        for m in S.metals:
            h.update(repr((m["priority"], m["order"])).encode() + box(m["prim"]))
        for m in S.lumped:
            h.update(repr((m["ny"], m["R"], m["caps"], np.asarray(m["lo"]).tolist(), np.asarray(m["hi"]).tolist())).encode())
        for m in S.excitations:
            h.update(repr((np.asarray(m["vec"]).tolist(), m["delay"], np.asarray(m["lo"]).tolist(), np.asarray(m["hi"]).tolist())).encode())
        for m in S.probes:
            h.update(repr((m["name"], m["p_type"], m["weight"], m["norm_dir"], np.asarray(m["start"]).tolist(), np.asarray(m["stop"]).tolist())).encode())
        nf = None if not S.nf2ff else (np.asarray(S.nf2ff["start"]).tolist(), np.asarray(S.nf2ff["stop"]).tolist(), repr(S.nf2ff.get("frequency")))
        h.update(repr((S.bc, S.pml_cells, S.f0, S.fc, S.timestep_factor, S.oversampling, nf, world, self.device,
                       None if self.nf2ff_freqs is None else np.asarray(self.nf2ff_freqs).tolist(),
                       np.asarray(S.probe_freqs).tolist(), repr(self.nf2ff_td), id(self.engine_factory))).encode())
        return h.hexdigest()

    # ---- the hot path ----
    def Run(self, sim_path, cleanup=False, setup_only=False, debug_material=False, debug_pec=False, debug_operator=False,
            debug_boxes=False, debug_csx=False, verbose=None, **kw):
        sim_path = os.path.abspath(str(sim_path))
        if cleanup and os.path.isdir(sim_path):
            shutil.rmtree(sim_path, ignore_errors=True)
        os.makedirs(sim_path, exist_ok=True)
        verbose = 0 if verbose is None else int(verbose)
        S = self._setup()
        rank, world, group = 0, 1, None
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and kw.get("distributed", True):
                rank, world = dist.get_rank(), dist.get_world_size()
        except Exception:
            pass
        log = (lambda msg: (print(msg), sys.stdout.flush())) if verbose else None
        key = self._scene_key(S, world)
        if self._prepared is not None and self._prepared[0] == key and S.nrts <= self._prepared[1].nrts_sized:
            # the scene was set up before (Run(..., setup_only=True)): ship the kept host operator to the device, reset the state
            # (buffers and the excitation signal were sized for at least this many steps)
            sim = self._prepared[1]
            sim.log = log or (lambda *a: None)
            sim.setup.nrts, sim.setup.end_criteria = S.nrts, S.end_criteria
            sim.restart()
        else:
            self._prepared = None
            sim = Simulation(S, device=self.device, rank=rank, world=world, group=group, engine_factory=self.engine_factory,
                             log=log, nf2ff_freqs=self.nf2ff_freqs, probe_freqs=S.probe_freqs,
                             fused_multi=getattr(self, "fused_multi", True), nf2ff_td=self.nf2ff_td)
            sim.prepare()
        self.sim = sim
        if verbose and rank == 0:
            nx, ny, nz = sim.nx, sim.ny, sim.nz_glob
            print(f"b200 FDTD engine: {nx}x{ny}x{nz} = {nx * ny * nz} cells, timestep {sim.dt:.4e} s, "
                  f"Nyquist {sim.nyquist} TS, excitation {sim.exc_len} TS, max {S.nrts} TS, "
                  f"{world} GPU(s); operator build {sim.prepare_s:.2f} s")
        if setup_only:
            self._prepared = (key, sim.keep_host_operator())
            return
        sim.run(verbose=verbose)
        res = sim.results
        res["device"] = self.device
        res["farfield_fn"] = self.farfield_fn
        res["stop_reason"] = sim.stop_reason
        self.results = res
        _registry.store(sim_path, res)
        if rank == 0:
            try:
                _ports.write_probe_files(sim_path, res)
            except OSError:
                pass
        if verbose and rank == 0:
            mc = sim.cells * sim.timesteps / max(sim.wall_s, 1e-9) / 1e6
            print(f"Time for {sim.timesteps} iterations with {sim.cells} cells : {sim.wall_s:.2f} sec")
            print(f"Speed: {mc:.1f} MCells/s ({sim.stop_reason})")
