"""Generate golden fixtures from the REFERENCE (run in the build container only; /root/reference is read-only).

1. closed_form.json  — values the reference's own code pins for config 1:
      design_patch_for_frequency (antenna_sim/physics.py:41-48), calculate_microstrip_width
      (antenna_sim/solver_fdtd_openems_microstrip.py:84-112), substrate kappa (…microstrip_3d.py:111)
2. trace_*.json      — the exact CSXCAD/openEMS call sequence the UNMODIFIED reference prepare functions emit
      (prepare_openems_microstrip_patch_3d, …microstrip_3d.py:19-196; prepare_openems_microstrip_multi_3d,
      …multi_3d.py:98-593), captured with recording stand-ins.  tests replay these against the shim, so the
      GPU box (which has no /root/reference) still drives the shim with the reference's own inputs.

3. run_prepared_*.npz — OUTPUTS of the UNMODIFIED live run functions (run_prepared_openems_microstrip_3d,
      …microstrip_3d.py:199-256; run_prepared_openems_microstrip_multi_3d, …multi_3d.py:596-663) executed here on the shim
      with the CPU oracle engine, at a frequency_hz that differs from the design frequency (2.40 GHz vs f0 = 2.45 GHz),
      time steps cut to keep the fixture run short.  The GPU box replays the same scene on the CUDA engine and must
      reproduce these dBi grids (tests/test_scene_parity.py::test_unmodified_run_prepared_outputs).

Usage:  python tests/golden/make_golden.py
"""
import json
import os
import sys
import types
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refload  # noqa: E402

C0 = 299792458.0
MUE0 = 4e-7 * np.pi
EPS0 = 1.0 / (MUE0 * C0 * C0)


def enc(v, ids):
    if isinstance(v, Rec):
        return {"ref": v._id}
    if isinstance(v, np.ndarray):
        return {"nd": v.tolist()}
    if isinstance(v, (np.floating, np.integer)):
        return v.item()
    if isinstance(v, (list, tuple)):
        return [enc(x, ids) for x in v]
    if isinstance(v, dict):
        return {k: enc(x, ids) for k, x in v.items()}
    return v


class Rec:
    """records every method call; each call returns a fresh recorder"""
    _log = []
    _n = 0

    def __init__(self, cls=None, args=(), kw=None):
        Rec._n += 1
        self._id = Rec._n
        if cls is not None:
            Rec._log.append({"new": cls, "id": self._id, "args": enc(list(args), None), "kw": enc(kw or {}, None)})

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)

        def call(*a, **k):
            out = Rec()
            Rec._log.append({"obj": self._id, "call": name, "args": enc(list(a), None), "kw": enc(k, None), "ret": out._id})
            return out
        return call


def install_mocks():
    Rec._log = []; Rec._n = 0
    csx = types.ModuleType("CSXCAD")
    csx.ContinuousStructure = lambda *a, **k: Rec("ContinuousStructure", a, k)
    csx.CSProperties = types.ModuleType("CSXCAD.CSProperties"); csx.CSPrimitives = types.ModuleType("CSXCAD.CSPrimitives")
    oe = types.ModuleType("openEMS")
    oe.openEMS = lambda *a, **k: Rec("openEMS", a, k)
    pc = types.ModuleType("openEMS.physical_constants")
    pc.C0, pc.EPS0, pc.MUE0, pc.Z0 = C0, EPS0, MUE0, float(np.sqrt(MUE0 / EPS0))
    oe.physical_constants = pc
    # the legacy backend also does `from openEMS import CSXCAD, nf2ff, ports, utilities, automesh`
    # (antenna_sim/solver_fdtd_openems.py:117-124)
    oe.CSXCAD = csx
    for sub in ("nf2ff", "ports", "utilities", "automesh"):
        setattr(oe, sub, types.ModuleType("openEMS." + sub))
    saved = {k: sys.modules.get(k) for k in ("CSXCAD", "openEMS", "openEMS.physical_constants")}
    sys.modules["CSXCAD"] = csx; sys.modules["openEMS"] = oe; sys.modules["openEMS.physical_constants"] = pc
    return saved


def restore(saved):
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


@dataclass
class PatchInstance:                         # antenna_sim/multi_patch_designer.py:18-28 (tkinter-free copy of the shape)
    name: str
    params: object
    center_x_m: float = 0.0
    center_y_m: float = 0.0
    center_z_m: float = 0.0
    rot_x_deg: float = 0.0
    rot_y_deg: float = 0.0
    rot_z_deg: float = 0.0
    feed_direction: object = None


def main():
    m = refload.load()
    P = m["models"].PatchAntennaParams.from_user_units(frequency_ghz=2.45, er=4.3, h_mm=1.6, loss_tangent=0.02, metal="copper")
    L, W, eeff = m["physics"].design_patch_for_frequency(P.frequency_hz, P.eps_r, P.h_m)
    fw = m["solver_fdtd_openems_microstrip"].calculate_microstrip_width(P.frequency_hz, P.eps_r, P.h_m)
    closed = dict(frequency_hz=P.frequency_hz, eps_r=P.eps_r, h_m=P.h_m, loss_tangent=P.loss_tangent,
                  patch_length_m=L, patch_width_m=W, eps_eff=eeff, feed_width_m=fw,
                  kappa=2 * np.pi * P.frequency_hz * EPS0 * P.eps_r * P.loss_tangent)
    json.dump(closed, open(os.path.join(HERE, "closed_form.json"), "w"), indent=1)
    FD = m["solver_fdtd_openems_microstrip"].FeedDirection
    cases = {
        "trace_single_pml8_q3": dict(boundary="PML_8", mesh_quality=3, feed_direction=FD.NEG_X),
        "trace_single_mur_q1": dict(boundary="MUR", mesh_quality=1, feed_direction=FD.NEG_X),
        "trace_single_mur_q2_posy": dict(boundary="MUR", mesh_quality=2, feed_direction=FD.POS_Y),
    }
    for name, kw in cases.items():
        saved = install_mocks()
        try:
            prep = m["solver_fdtd_openems_microstrip_3d"].prepare_openems_microstrip_patch_3d(
                P, dll_dir=refload.DLL_DIR, theta_step_deg=2.0, phi_step_deg=5.0, **kw)
        finally:
            restore(saved)
        assert prep.ok, prep.message
        out = dict(case=name, kwargs={k: str(v) for k, v in kw.items()}, log=Rec._log, fdtd=prep.FDTD._id, nf=prep.nf._id,
                   theta=prep.theta.tolist(), phi=prep.phi.tolist(), nf_center=prep.nf_center.tolist())
        json.dump(out, open(os.path.join(HERE, name + ".json"), "w"))
        print(name, len(Rec._log), "calls")
    # sibling callers: tutorial-exact scene, E/H-cut microstrip variant, legacy int-BC backend
    sib = refload.load(("solver_fdtd_openems",))
    for name, fn in (("trace_fixed_tutorial", lambda: m["solver_fdtd_openems_fixed"].prepare_openems_patch_fixed(P, dll_dir=refload.DLL_DIR)),
                     ("trace_microstrip_cuts", lambda: m["solver_fdtd_openems_microstrip"].prepare_openems_microstrip_patch(P, dll_dir=refload.DLL_DIR)),
                     ("trace_legacy_intbc", lambda: sib["solver_fdtd_openems"].prepare_openems_patch(P, dll_dir=refload.DLL_DIR))):
        saved = install_mocks()
        try:
            prep = fn()
        finally:
            restore(saved)
        assert prep.ok, prep.message
        out = dict(case=name, log=Rec._log, fdtd=prep.FDTD._id, nf=prep.nf._id,
                   theta=np.asarray(prep.theta).tolist(), phi=np.asarray(prep.phi).tolist(),
                   nf_center=(np.asarray(prep.nf_center).tolist() if getattr(prep, "nf_center", None) is not None else [0.0, 0.0, 0.0]))
        json.dump(out, open(os.path.join(HERE, name + ".json"), "w"))
        print(name, len(Rec._log), "calls")
    # multi-patch: 2 elements, quality 4, copper 35 um
    saved = install_mocks()
    try:
        patches = [PatchInstance("P1", P, center_x_m=-0.035, feed_direction=FD.NEG_X),
                   PatchInstance("P2", P, center_x_m=0.035, rot_z_deg=90.0, feed_direction=FD.NEG_X)]
        prep = m["solver_fdtd_openems_microstrip_multi_3d"].prepare_openems_microstrip_multi_3d(
            patches, dll_dir=refload.DLL_DIR, boundary="MUR", mesh_quality=2, theta_step_deg=5.0, phi_step_deg=15.0)
    finally:
        restore(saved)
    assert prep.ok, prep.message
    out = dict(case="trace_multi2_mur_q2", log=Rec._log, fdtd=prep.FDTD._id, nf=prep.nf._id,
               theta=prep.theta.tolist(), phi=prep.phi.tolist(), nf_center=np.asarray(prep.nf_center).tolist())
    json.dump(out, open(os.path.join(HERE, "trace_multi2_mur_q2.json"), "w"))
    print("trace_multi2_mur_q2", len(Rec._log), "calls")
    # BASELINE.json configs[2]: 4x4 array of PatchInstances (multi_patch_designer.py:18-28), pitch 60 mm (~ lambda0/2), all
    # elements fed from -x; quality 1 + MUR for the parity tests, quality 2 + PML_8 as the base of the ~1 B-cell bench mesh
    for name, bc, q in (("trace_array16_mur_q1", "MUR", 1), ("trace_array16_pml8_q2", "PML_8", 2)):
        saved = install_mocks()
        try:
            pitch = 0.060
            patches = [PatchInstance(f"P{r}{c}", P, center_x_m=(c - 1.5) * pitch, center_y_m=(r - 1.5) * pitch, feed_direction=FD.NEG_X)
                       for r in range(4) for c in range(4)]
            prep = m["solver_fdtd_openems_microstrip_multi_3d"].prepare_openems_microstrip_multi_3d(
                patches, dll_dir=refload.DLL_DIR, boundary=bc, mesh_quality=q, theta_step_deg=5.0, phi_step_deg=15.0)
        finally:
            restore(saved)
        assert prep.ok, prep.message
        out = dict(case=name, log=Rec._log, fdtd=prep.FDTD._id, nf=prep.nf._id,
                   theta=prep.theta.tolist(), phi=prep.phi.tolist(), nf_center=np.asarray(prep.nf_center).tolist())
        json.dump(out, open(os.path.join(HERE, name + ".json"), "w"))
        print(name, len(Rec._log), "calls")


RUN_PREPARED_F = 2.40e9
RUN_PREPARED_STEPS = {"single_mur_q1": 3000, "multi2_mur_q2": 1200}


def run_prepared_outputs():
    """the unmodified run_prepared_* functions on the real shim + oracle engine -> tests/golden/run_prepared_*.npz"""
    root = os.path.dirname(os.path.dirname(HERE))
    for p_ in (root, os.path.join(root, "fdtd-solver-antennas_b200")):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    import scenes
    scenes.use_oracle_engine(threads=os.cpu_count() or 4)
    m = refload.load()
    P = m["models"].PatchAntennaParams.from_user_units(frequency_ghz=2.45, er=4.3, h_mm=1.6, loss_tangent=0.02, metal="copper")
    FD = m["solver_fdtd_openems_microstrip"].FeedDirection
    s3, sm = m["solver_fdtd_openems_microstrip_3d"], m["solver_fdtd_openems_microstrip_multi_3d"]
    prep = s3.prepare_openems_microstrip_patch_3d(P, dll_dir=refload.DLL_DIR, boundary="MUR", mesh_quality=1, feed_direction=FD.NEG_X,
                                                  theta_step_deg=2.0, phi_step_deg=5.0)
    assert prep.ok, prep.message
    prep.FDTD.SetNumberOfTimeSteps(RUN_PREPARED_STEPS["single_mur_q1"])
    res = s3.run_prepared_openems_microstrip_3d(prep, frequency_hz=RUN_PREPARED_F, verbose=0)
    assert res.ok, res.message
    np.savez_compressed(os.path.join(HERE, "run_prepared_single_mur_q1.npz"), intensity=res.intensity.astype(np.float32),
                        theta=res.theta, phi=res.phi, frequency_hz=RUN_PREPARED_F, steps=prep.FDTD.sim.timesteps)
    print("run_prepared_single_mur_q1", res.intensity.shape, float(res.intensity.max()), prep.FDTD.sim.timesteps, prep.FDTD.sim.stop_reason)
    patches = [PatchInstance("P1", P, center_x_m=-0.035, feed_direction=FD.NEG_X),
               PatchInstance("P2", P, center_x_m=0.035, rot_z_deg=90.0, feed_direction=FD.NEG_X)]
    prep = sm.prepare_openems_microstrip_multi_3d(patches, dll_dir=refload.DLL_DIR, boundary="MUR", mesh_quality=2,
                                                  theta_step_deg=5.0, phi_step_deg=15.0)
    assert prep.ok, prep.message
    prep.FDTD.SetNumberOfTimeSteps(RUN_PREPARED_STEPS["multi2_mur_q2"])
    prep.FDTD.SetEndCriteria(1e-30)
    res = sm.run_prepared_openems_microstrip_multi_3d(prep, frequency_hz=RUN_PREPARED_F, verbose=0)
    assert res.ok, res.message
    np.savez_compressed(os.path.join(HERE, "run_prepared_multi2_mur_q2.npz"), intensity=res.intensity.astype(np.float32),
                        theta=res.theta, phi=res.phi, frequency_hz=RUN_PREPARED_F, steps=prep.FDTD.sim.timesteps)
    print("run_prepared_multi2_mur_q2", res.intensity.shape, float(res.intensity.max()), prep.FDTD.sim.timesteps)
    scenes.use_cuda_engine()


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] != "outputs":
        main()
    run_prepared_outputs()
