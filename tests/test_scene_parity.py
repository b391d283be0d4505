"""End-to-end parity, CUDA engine vs CPU oracle, through the drop-in API on identical meshes and excitations.

North-star tolerances (BASELINE.json): S11 within 0.2 dB over band, resonant frequency within 0.1 %, far-field
gain within 0.1 dBi (fp32).  Both engines consume the same operator, so the fields are in fact bit-identical;
the outputs differ only by fp32-vs-fp64 accumulation in the probe sums and running DFTs.
"""
import os

import numpy as np
import pytest

import replay
import scenes

pytestmark = pytest.mark.gpu

S11_TOL_DB = 0.2
FRES_TOL = 1e-3
GAIN_TOL_DB = 0.1


def _run_both(build, nrts, tag):
    out = {}
    for eng in ("oracle", "cuda"):
        (scenes.use_oracle_engine(threads=os.cpu_count() or 4) if eng == "oracle" else scenes.use_cuda_engine())
        F, nf, port, theta, phi, center = build()
        F.SetNumberOfTimeSteps(nrts)
        path = scenes.tmp_sim_path(f"{tag}_{eng}")
        F.Run(path, cleanup=True)
        f0 = F.exc[1]
        f, s11, zin = replay.s11_db(port, path, f0)
        dbi, dmax = replay.reference_postprocess(nf, path, f0, theta, phi, center)
        E = F.sim.engine
        volt = E.volt.cpu().numpy() if hasattr(E.volt, "cpu") else E.volt.copy()
        out[eng] = dict(f=f, s11=s11, zin=zin, dbi=dbi, dmax=dmax, ts=F.sim.timesteps, volt=volt, stop=F.sim.stop_reason)
    scenes.use_cuda_engine()
    return out


def _check(out):
    o, c = out["oracle"], out["cuda"]
    assert o["ts"] == c["ts"], "engines stopped at different time steps"
    assert np.array_equal(o["volt"].view(np.uint32), c["volt"].view(np.uint32)), "final E field is not bit-identical"
    assert np.abs(o["s11"] - c["s11"]).max() <= S11_TOL_DB, f"S11 differs by {np.abs(o['s11'] - c['s11']).max()} dB"
    fo, fc = o["f"][np.argmin(o["s11"])], c["f"][np.argmin(c["s11"])]
    assert abs(fo - fc) <= FRES_TOL * fo
    assert abs(10 * np.log10(o["dmax"]) - 10 * np.log10(c["dmax"])) <= GAIN_TOL_DB
    sel = o["dbi"] > o["dbi"].max() - 30.0
    assert np.abs(o["dbi"][sel] - c["dbi"][sel]).max() <= GAIN_TOL_DB
    assert np.abs(o["zin"] - c["zin"]).max() <= 1e-3 * np.abs(o["zin"]).max()


def test_dipole_pml():
    def build():
        F, nf, port = scenes.dipole("PML_8", cells=(40, 40, 48), nrts=3000, end=1e-6)
        return F, nf, port, np.arange(0.0, 181.0, 5.0), np.arange(0.0, 361.0, 45.0), np.zeros(3)
    out = _run_both(build, 3000, "dip")
    _check(out)
    assert 1.45 < out["cuda"]["dmax"] < 1.62


@pytest.mark.parametrize("case,nrts", [("trace_single_mur_q1", 6000), ("trace_single_pml8_q3", 3000), ("trace_multi2_mur_q2", 1200),
                                       ("trace_array16_mur_q1", 1300),
                                       ("trace_fixed_tutorial", 3000), ("trace_microstrip_cuts", 2000), ("trace_legacy_intbc", 1500)])
def test_reference_scenes(case, nrts):
    """the reference's own scenes (recorded call traces of the unmodified prepare functions), both engines"""
    def build():
        R = replay.replay(case)
        th, ph = R["theta"], R["phi"]
        if th.max() <= 2 * np.pi + 1e-9 and ph.max() <= 2 * np.pi + 1e-9 and len(th) > 8:
            th, ph = np.rad2deg(th), np.rad2deg(ph)          # the legacy backend keeps radians (solver_fdtd_openems.py)
        th = th[::5] if len(th) > 20 else th
        ph = ph[::9] if len(ph) > 20 else ph
        return R["FDTD"], R["nf"], R["FDTD"].ports[0], th, ph, R["nf_center"]
    out = _run_both(build, nrts, case)
    _check(out)


@pytest.mark.parametrize("case,trace", [("single_mur_q1", "trace_single_mur_q1"), ("multi2_mur_q2", "trace_multi2_mur_q2")])
def test_unmodified_run_prepared_outputs(case, trace):
    """SURVEY.md §8 a2/a3 on CUDA: the committed fixtures are the dBi grids returned by the UNMODIFIED
    run_prepared_openems_microstrip_3d / …multi_3d (tests/golden/make_golden.py, CPU oracle engine, frequency_hz = 2.40 GHz,
    which is not the registered design frequency).  The same scene (recorded call trace of the unmodified prepare) runs
    here on the CUDA engine; far field at 2.40 GHz from the stored face samples; gain grid within 0.1 dB."""
    G = np.load(os.path.join(replay.GOLDEN, f"run_prepared_{case}.npz"))
    scenes.use_cuda_engine()
    R = replay.replay(trace)
    F = R["FDTD"]
    F.SetNumberOfTimeSteps(int(G["steps"]))
    if case.startswith("multi"):
        F.SetEndCriteria(1e-30)
    path = scenes.tmp_sim_path("rp_" + case)
    F.Run(path, cleanup=True)
    assert F.sim.timesteps == int(G["steps"])
    assert F.sim.td_store
    dbi, dmax = replay.reference_postprocess(R["nf"], path, float(G["frequency_hz"]), R["theta"], R["phi"], R["nf_center"])
    gold = G["intensity"].astype(np.float64)
    assert dbi.shape == gold.shape
    assert np.allclose(np.deg2rad(R["theta"]), G["theta"]) and np.allclose(np.deg2rad(R["phi"]), G["phi"])
    sel = gold > gold.max() - 30.0
    assert np.abs(dbi[sel] - gold[sel]).max() <= GAIN_TOL_DB, np.abs(dbi[sel] - gold[sel]).max()
    assert abs(dbi.max() - gold.max()) <= GAIN_TOL_DB


def test_config4_broadband_multi_frequency():
    """BASELINE.json configs[3]: broadband 1-6 GHz pulse (f0 3.5 GHz, fc 2.5 GHz), running DFT of the port V/I at 501
    points and of the NF2FF box at 11 frequencies, all on the device; compared with the oracle's double-precision DFTs"""
    from b200fdtd import scenes as pscenes
    nf_freqs = np.linspace(1e9, 6e9, 11)
    pf = np.linspace(1e9, 6e9, 501)
    out = {}
    for eng in ("oracle", "cuda"):
        (scenes.use_oracle_engine(threads=os.cpu_count() or 4) if eng == "oracle" else scenes.use_cuda_engine())
        F, nf, port = pscenes.patch_scene(mesh_res_mm=5.0, boundary="PML_8", f0=3.5e9, fc=2.5e9, nrts=2500, end_criteria=1e-9,
                                          nf2ff_freqs=nf_freqs)
        F.port_dft_freqs = pf
        path = scenes.tmp_sim_path(f"cfg4_{eng}")
        F.Run(path, cleanup=True)
        port.CalcPort(path, pf)                       # registered grid -> device running DFT on the CUDA engine
        s11 = 20 * np.log10(np.abs(port.uf_ref / port.uf_inc))
        theta, phi = np.arange(0.0, 181.0, 10.0), np.array([0.0, 90.0])
        res = nf.CalcNF2FF(path, nf_freqs, theta, phi, center=[0, 0, 0.8e-3])
        out[eng] = dict(s11=s11, dmax=np.array(res.Dmax), e=np.array(res.E_norm), used_dft=F.results["probes"]["port_ut_1"]["dft"] is not None)
    scenes.use_cuda_engine()
    o, c = out["oracle"], out["cuda"]
    assert c["used_dft"]
    assert np.abs(o["s11"] - c["s11"]).max() <= S11_TOL_DB
    assert np.abs(10 * np.log10(o["dmax"]) - 10 * np.log10(c["dmax"])).max() <= GAIN_TOL_DB
    for q in range(len(nf_freqs)):
        eo, ec = o["e"][q], c["e"][q]
        sel = eo > eo.max() * 10 ** (-30 / 20)
        assert np.abs(20 * np.log10(ec[sel] / eo[sel])).max() <= GAIN_TOL_DB


def test_compressed_operator_host_round_trip():
    """export the operator in its compressed host form, wipe the device arrays, reload: bit-identical arrays,
    no row demoted, and a run from the reloaded operator equals the oracle"""
    import torch
    from b200fdtd import scenes as pscenes
    from b200fdtd.simulation import Simulation
    scenes.use_cuda_engine()
    F, nf, port = pscenes.patch_scene(mesh_res_mm=4.0, boundary="PML_8", nrts=400, end_criteria=1e-12)
    S = F._setup()
    sim = Simulation(S, device=0, nf2ff_freqs=F.nf2ff_freqs, probe_freqs=S.probe_freqs).prepare()
    E = sim.engine
    assert sim.compression[0][0] > 0 and sim.compression[0][1] == 0 and sim.compression[1][1] == 0
    op = sim.export_operator(pin=True)
    full_bytes = 4 * E.vv.numel() * 4
    assert sim.operator_nbytes(op) < 0.05 * full_bytes
    before = [t.clone() for t in (E.vv, E.vi, E.ii, E.iv)]
    for t in (E.vv, E.vi, E.ii, E.iv):
        t.zero_()
    res = sim.load_operator(op)
    assert res[0] == sim.compression[0][:2] and res[1] == sim.compression[1][:2]
    for a, b in zip(before, (E.vv, E.vi, E.ii, E.iv)):
        assert torch.equal(a, b)
    E.run(60)
    from oracle.fdtd_ref import RefEngine
    sim_o = Simulation(S, device=0, engine_factory=lambda nx, ny, nz, px, dev: RefEngine(nx, ny, nz, px, threads=os.cpu_count() or 4),
                       nf2ff_freqs=F.nf2ff_freqs, probe_freqs=S.probe_freqs).prepare()
    sim_o.engine.run(60)
    nz = sim.nz
    assert np.array_equal(E.volt.cpu().numpy()[:, 1:nz + 1].view(np.uint32), sim_o.engine.volt[:, 1:nz + 1].view(np.uint32))


def test_config1_full_length_run_to_end_criteria():
    """BASELINE.json configs[0]: the reference's own single-patch scene (unmodified prepare, PML_8, quality 3) run to its
    end criterion (NrTS 30000, EndCriteria 1e-4) on both engines: same stopping step, bit-identical final field, S11 /
    resonance / gain within the north-star tolerances"""
    def build():
        R = replay.replay("trace_single_pml8_q3")
        return R["FDTD"], R["nf"], R["FDTD"].ports[0], R["theta"][::5], R["phi"][::9], R["nf_center"]
    out = _run_both(build, 30000, "cfg1_full")
    _check(out)
    assert out["cuda"]["stop"] == "EndCriteria" and out["cuda"]["ts"] < 30000


def test_config2_size_parity_100m_cells():
    """BASELINE.json configs[1] at full size: the patch scene meshed at ~106 M cells with PML_8 (20 z-chunks of the fused
    launch, 5 x-segments, byte offsets beyond 2^31, row-compressed operator built on the GPU): one sampling interval by
    graph replay of the fused H->E path plus a few eager steps on the CUDA engine, the same steps on the CPU oracle from
    the same operator -> fields bit-identical, probe samples and NF2FF accumulators within fp32 accumulation tolerance"""
    import torch
    from b200fdtd import scenes as pscenes
    from b200fdtd.simulation import Simulation
    from oracle.fdtd_ref import RefEngine
    if torch.cuda.mem_get_info()[0] < 40 * 2 ** 30:
        pytest.skip("needs 40 GB of free HBM")
    scenes.use_cuda_engine()
    F, nf, port = pscenes.patch_scene(target_cells=100e6, boundary="PML_8", f0=2.5e9, fc=1.5e9, nrts=10 ** 6, end_criteria=1e-12,
                                      nf2ff_freqs=[2.45e9])
    S = F._setup()
    dev = torch.device("cuda", 0)
    sim = Simulation(S, device=0, nf2ff_freqs=F.nf2ff_freqs, probe_freqs=S.probe_freqs, build_device=dev, nf2ff_td=False).prepare()
    assert sim.cells > 100e6
    n = sim.interval + 3
    E = sim.engine
    # seeded noise in every cell of both fields (the pulse alone would leave most of the grid at zero after so few steps)
    g = torch.Generator(device=dev).manual_seed(0)
    for f in (E.volt, E.curr):
        f[:, 1:sim.nz + 1, :, :sim.nx] = 1e-3 * torch.randn((3, sim.nz, sim.ny, sim.nx), generator=g, device=dev)
    v0, c0 = E.volt.cpu().numpy(), E.curr.cpu().numpy()
    E.run(n, use_graph=True)
    assert E.he_active and E.plan_info()[2] == 0
    sim_o = Simulation(S, device=0, engine_factory=lambda nx, ny, nz, px, d: RefEngine(nx, ny, nz, px, threads=os.cpu_count() or 8),
                       nf2ff_freqs=F.nf2ff_freqs, probe_freqs=S.probe_freqs, build_device=dev, nf2ff_td=False).prepare()
    O = sim_o.engine
    assert (sim_o.nx, sim_o.ny, sim_o.nz, sim_o.px) == (sim.nx, sim.ny, sim.nz, sim.px) and sim_o.dt == sim.dt
    O.volt[...] = v0; O.curr[...] = c0
    del v0, c0
    O.run(n)
    nz = sim.nz
    for name, g, o in (("volt", E.volt, O.volt), ("curr", E.curr, O.curr)):
        for c in range(3):
            gh = g[c, 1:nz + 1].cpu().numpy()
            assert np.array_equal(gh.view(np.uint32), o[c, 1:nz + 1].view(np.uint32)), f"{name}[{c}] differs at 106 M cells"
            assert np.abs(gh).max() > 0 or c < 2
            del gh
    ns = n // sim.interval
    s_g = E.series.cpu().numpy()[:, :ns].astype(np.float64)
    assert np.abs(s_g - O.series[:, :ns]).max() <= 1e-5 * np.abs(O.series[:, :ns]).max()
    for fa_g, fa_r in zip(E.face_acc, O.face_acc):
        assert np.abs(fa_g.cpu().numpy() - fa_r).max() <= 1e-5 * max(np.abs(fa_r).max(), 1e-300)
