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
