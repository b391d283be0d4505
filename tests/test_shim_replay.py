"""The shim accepts the reference's exact call sequences (recorded fixtures) and, where the reference tree is
present, the UNMODIFIED reference prepare functions themselves."""
import json
import os

import numpy as np
import pytest

import refload
import replay

GOLDEN = replay.GOLDEN
CASES = ["trace_single_pml8_q3", "trace_single_mur_q1", "trace_single_mur_q2_posy", "trace_multi2_mur_q2",
         "trace_fixed_tutorial", "trace_microstrip_cuts", "trace_legacy_intbc"]


@pytest.mark.parametrize("case", CASES)
def test_replay_builds_a_valid_scene(case):
    R = replay.replay(case)
    F = R["FDTD"]
    S = F._setup()
    n = [len(l) for l in S.lines]
    assert all(m > 12 for m in n)
    for l in S.lines:
        d = np.diff(l)
        assert (d > 0).all()
    assert len(F.ports) >= 1 and len(F.nf2ff_boxes) == 1
    assert len(S.lumped) == len(F.ports) and len(S.excitations) == len(F.ports)
    assert len(S.probes) == 2 * len(F.ports)
    assert len(S.metals) >= 2
    # every port coordinate sits exactly on mesh lines (edges2grid / explicit AddLine in the reference)
    unit = F.GetCSX().GetGrid().GetDeltaUnit()
    for p in F.ports:
        for a in range(3):
            for v in (p.start[a], p.stop[a]):
                assert np.min(np.abs(S.lines[a] - v * unit)) < 1e-12


def test_closed_form_pins_in_the_single_patch_trace():
    """patch/feed geometry in the recorded calls equals the reference's closed-form design values (App. B)"""
    cf = json.load(open(os.path.join(GOLDEN, "closed_form.json")))
    assert abs(cf["patch_length_m"] * 1e3 - 29.138326) < 1e-5
    assert abs(cf["patch_width_m"] * 1e3 - 37.583886) < 1e-5
    assert abs(cf["eps_eff"] - 3.992370) < 1e-6
    assert abs(cf["feed_width_m"] * 1e3 - 3.114396) < 1e-5
    assert abs(cf["kappa"] - 0.0117218) < 1e-6
    R = replay.replay("trace_single_pml8_q3")
    csx = R["FDTD"].GetCSX()
    patch = csx.GetPropertiesByName("patch")[0].primitives[0]
    assert np.allclose(patch.stop - patch.start, [cf["patch_width_m"] * 1e3, cf["patch_length_m"] * 1e3, 0.0], rtol=1e-12)
    feed = csx.GetPropertiesByName("feed_line")[0].primitives[0]
    assert np.isclose(feed.stop[1] - feed.start[1], cf["feed_width_m"] * 1e3, rtol=1e-12)
    sub = csx.GetPropertiesByName("substrate")[0]
    assert np.isclose(sub.props["kappa"], cf["kappa"], rtol=1e-12) and sub.props["epsilon"] == 4.3
    S = R["FDTD"]._setup()
    assert S.bc == [3] * 6 and S.pml_cells == [8] * 6 and S.nrts == 30000 and S.end_criteria == 1e-4
    assert S.f0 == 2.45e9 and S.fc == 1.225e9
    assert len(R["theta"]) == 91 and len(R["phi"]) == 73


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("boundary,quality", [("PML_8", 3), ("MUR", 1)])
def test_unmodified_reference_prepare_runs_on_the_shim(boundary, quality):
    m = refload.load()
    P = m["models"].PatchAntennaParams.from_user_units(frequency_ghz=2.45, er=4.3, h_mm=1.6, loss_tangent=0.02, metal="copper")
    prep = m["solver_fdtd_openems_microstrip_3d"].prepare_openems_microstrip_patch_3d(
        P, dll_dir=refload.DLL_DIR, boundary=boundary, mesh_quality=quality)
    assert prep.ok, prep.message
    case = "trace_single_pml8_q3" if boundary == "PML_8" else "trace_single_mur_q1"
    R = replay.replay(case)
    Sa, Sb = prep.FDTD._setup(), R["FDTD"]._setup()
    for la, lb in zip(Sa.lines, Sb.lines):
        assert np.array_equal(la, lb)
    assert Sa.bc == Sb.bc and len(Sa.metals) == len(Sb.metals)
    assert np.array_equal(prep.theta, R["theta"]) and np.array_equal(prep.phi, R["phi"])


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_unmodified_reference_multi_prepare_runs_on_the_shim():
    from dataclasses import dataclass
    m = refload.load()
    P = m["models"].PatchAntennaParams.from_user_units(frequency_ghz=2.45, er=4.3, h_mm=1.6, loss_tangent=0.02, metal="copper")
    FD = m["solver_fdtd_openems_microstrip"].FeedDirection

    @dataclass
    class PI:
        name: str
        params: object
        center_x_m: float = 0.0
        center_y_m: float = 0.0
        center_z_m: float = 0.0
        rot_x_deg: float = 0.0
        rot_y_deg: float = 0.0
        rot_z_deg: float = 0.0
        feed_direction: object = None
    patches = [PI("P1", P, center_x_m=-0.035, feed_direction=FD.NEG_X), PI("P2", P, center_x_m=0.035, rot_z_deg=90.0, feed_direction=FD.NEG_X)]
    prep = m["solver_fdtd_openems_microstrip_multi_3d"].prepare_openems_microstrip_multi_3d(
        patches, dll_dir=refload.DLL_DIR, boundary="MUR", mesh_quality=2, theta_step_deg=5.0, phi_step_deg=15.0)
    assert prep.ok, prep.message
    R = replay.replay("trace_multi2_mur_q2")
    Sa, Sb = prep.FDTD._setup(), R["FDTD"]._setup()
    for la, lb in zip(Sa.lines, Sb.lines):
        assert np.array_equal(la, lb)
    assert len(prep.FDTD.ports) == 2


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_sibling_callers_run_unmodified_end_to_end():
    """tutorial-exact scene, E/H-cut microstrip variant and the legacy int-BC backend run unmodified
    (prepare AND run, incl. their own CalcNF2FF post-processing) on the shim; engine = oracle here, CUDA on the GPU box"""
    import scenes
    scenes.use_oracle_engine(threads=4)
    try:
        m = refload.load(("models", "physics", "solver_fdtd_openems_fixed", "solver_fdtd_openems_microstrip", "solver_fdtd_openems"))
        P = m["models"].PatchAntennaParams.from_user_units(frequency_ghz=2.45, er=4.3, h_mm=1.6, loss_tangent=0.02, metal="copper")
        fx, ms, lg = m["solver_fdtd_openems_fixed"], m["solver_fdtd_openems_microstrip"], m["solver_fdtd_openems"]
        assert fx.probe_openems_fixed(refload.DLL_DIR).ok
        assert ms.probe_openems_microstrip(refload.DLL_DIR).ok
        tmp = scenes.tmp_sim_path("sib")
        for prep_fn, run_fn in ((lambda: fx.prepare_openems_patch_fixed(P, dll_dir=refload.DLL_DIR, work_dir=tmp + "_fx"), fx.run_prepared_openems_fixed),
                                (lambda: ms.prepare_openems_microstrip_patch(P, dll_dir=refload.DLL_DIR, work_dir=tmp + "_ms"), ms.run_prepared_openems_microstrip),
                                (lambda: lg.prepare_openems_patch(P, dll_dir=refload.DLL_DIR, work_dir=tmp + "_lg"), lg.run_prepared_openems)):
            prep = prep_fn()
            assert prep.ok, prep.message
            prep.FDTD.SetNumberOfTimeSteps(400)
            res = run_fn(prep, frequency_hz=P.frequency_hz, verbose=0)
            assert res.ok, res.message
            assert res.intensity is not None and np.all(np.isfinite(np.asarray(res.intensity)))
    finally:
        scenes.use_cuda_engine()


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_live_run_prepared_functions_run_unmodified(tmp_path, monkeypatch):
    """the two LIVE run functions (run_prepared_openems_microstrip_3d, …microstrip_3d.py:199-256;
    run_prepared_openems_microstrip_multi_3d, …multi_3d.py:596-663) executed unmodified on the shim, at the design
    frequency and at a frequency_hz that was never registered (2.40 GHz; served from the stored NF2FF face samples).
    They must return ok=True and exactly what the test-side restatement of their loop (replay.reference_postprocess,
    used by the GPU parity tests) returns; the committed fixtures run_prepared_*.npz are outputs of these same calls."""
    import scenes
    from dataclasses import dataclass
    monkeypatch.chdir(tmp_path)                       # the reference creates its sim_path relative to the cwd
    scenes.use_oracle_engine(threads=os.cpu_count() or 4)
    try:
        m = refload.load()
        P = m["models"].PatchAntennaParams.from_user_units(frequency_ghz=2.45, er=4.3, h_mm=1.6, loss_tangent=0.02, metal="copper")
        FD = m["solver_fdtd_openems_microstrip"].FeedDirection
        s3, sm = m["solver_fdtd_openems_microstrip_3d"], m["solver_fdtd_openems_microstrip_multi_3d"]
        prep = s3.prepare_openems_microstrip_patch_3d(P, dll_dir=refload.DLL_DIR, boundary="MUR", mesh_quality=1, feed_direction=FD.NEG_X,
                                                      theta_step_deg=10.0, phi_step_deg=45.0)
        assert prep.ok, prep.message
        prep.FDTD.SetNumberOfTimeSteps(1500)
        for f in (P.frequency_hz, 2.40e9):
            res = s3.run_prepared_openems_microstrip_3d(prep, frequency_hz=f, verbose=0)
            assert res.ok, res.message
            assert res.is_dBi and res.intensity.shape == (len(prep.theta), len(prep.phi)) and np.isfinite(res.intensity).all()
            dbi, dmax = replay.reference_postprocess(prep.nf, res.sim_path, f, prep.theta, prep.phi, prep.nf_center)
            assert np.array_equal(dbi, res.intensity)
            assert np.allclose(res.theta, np.deg2rad(prep.theta)) and np.allclose(res.phi, np.deg2rad(prep.phi))

        @dataclass
        class PI:
            name: str
            params: object
            center_x_m: float = 0.0
            center_y_m: float = 0.0
            center_z_m: float = 0.0
            rot_x_deg: float = 0.0
            rot_y_deg: float = 0.0
            rot_z_deg: float = 0.0
            feed_direction: object = None
        patches = [PI("P1", P, center_x_m=-0.035, feed_direction=FD.NEG_X), PI("P2", P, center_x_m=0.035, rot_z_deg=90.0, feed_direction=FD.NEG_X)]
        prep = sm.prepare_openems_microstrip_multi_3d(patches, dll_dir=refload.DLL_DIR, boundary="MUR", mesh_quality=1,
                                                      theta_step_deg=15.0, phi_step_deg=45.0)
        assert prep.ok, prep.message
        prep.FDTD.SetNumberOfTimeSteps(500); prep.FDTD.SetEndCriteria(1e-30)
        res = sm.run_prepared_openems_microstrip_multi_3d(prep, frequency_hz=2.40e9, verbose=0)
        assert res.ok, res.message
        dbi, dmax = replay.reference_postprocess(prep.nf, res.sim_path, 2.40e9, prep.theta, prep.phi, prep.nf_center)
        assert np.array_equal(dbi, res.intensity)
    finally:
        scenes.use_cuda_engine()


def test_setup_only_keeps_the_prepared_scene_and_a_changed_scene_is_rebuilt():
    """FDTD.Run(sim_path, setup_only=True) (openEMS's own flag) keeps the prepared scene; a later Run of the SAME scene
    restarts it from zero fields and reproduces the first run; a shorter NrTS reuses it, a changed scene does not"""
    import scenes
    scenes.use_oracle_engine(threads=4)
    try:
        F, nf, port = scenes.dipole("PML_8", cells=(20, 20, 24), nrts=240, end=1e-12)
        p = scenes.tmp_sim_path("reuse")
        F.Run(p, cleanup=True)
        first = F.results["probes"]["port_ut_1"]["val"].copy()
        acc0 = [a.copy() for a in F.results["nf2ff"]["acc"]]
        F.Run(p, setup_only=True)
        sim = F._prepared[1]
        for _ in range(2):
            F.Run(p)
            assert F.sim is sim, "the prepared scene was not reused"
            assert np.array_equal(F.results["probes"]["port_ut_1"]["val"], first)
            for a, b in zip(F.results["nf2ff"]["acc"], acc0):
                assert np.array_equal(a, b)
        F.SetNumberOfTimeSteps(120)
        F.Run(p)
        assert F.sim is sim and F.sim.timesteps == 120
        assert np.array_equal(F.results["probes"]["port_ut_1"]["val"], first[:len(F.results["probes"]["port_ut_1"]["val"])])
        F.GetCSX().AddMetal("extra").AddBox([1.0, 1.0, 1.0], [4.0, 4.0, 1.0], priority=10)     # the scene changes
        F.Run(p)
        assert F.sim is not sim, "a changed scene must be prepared again"
    finally:
        scenes.use_cuda_engine()


def test_a_z_slab_builds_the_rows_of_the_whole_grid_bit_for_bit():
    """the float32 operator is defined row by row (operator.py:RowFactor), so what a z-slab rank builds for its planes is
    bit-identical to the single-GPU build, whatever the chunking and whichever rows share an x-vector (16-element array:
    translated boxes, 16 ports, thick copper)"""
    import torch
    from b200fdtd.operator import OperatorBuilder
    S = replay.replay("trace_array16_mur_q1")["FDTD"]._setup()
    nz = len(S.lines[2])
    px = (len(S.lines[0]) + 31) // 32 * 32
    B = OperatorBuilder(S, device=torch.device("cpu"), k_nodes=(0, nz))
    dt = B.estimate_timestep()
    full = [t.reshape(3, nz + 2, -1, px) for t in B.coefficients(0, nz, px, dt)]
    cuts = [0, nz // 3, nz // 2 + 1, nz]
    dts = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        Bs = OperatorBuilder(S, device=torch.device("cpu"), k_nodes=(a, b))
        dts.append(Bs.estimate_timestep())
        part = [t.reshape(3, b - a + 2, -1, px) for t in Bs.coefficients(a, b, px, dt, chunk=7)]
        for name, f, p in zip(("vv", "vi", "ii", "iv"), full, part):
            assert torch.equal(f[:, a + 1:b + 1].view(torch.int32), p[:, 1:-1].view(torch.int32)), (name, a, b)
    assert min(dts) == dt                      # the global time step is the minimum over the slabs
