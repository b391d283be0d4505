"""CalcNF2FF at a frequency that was NOT registered for the running DFT (the reference passes the caller's frequency_hz:
antenna_sim/solver_fdtd_openems_microstrip_3d.py:225; openEMS transforms its time-domain HDF5 dumps at any frequency).
The engine keeps the face samples in memory when they fit and transforms them on demand."""
import os

import numpy as np
import pytest

import scenes


def _run(engine, nf2ff_freqs, tag, td=None):
    (scenes.use_oracle_engine(threads=os.cpu_count() or 4) if engine == "oracle" else scenes.use_cuda_engine())
    F, nf, port = scenes.dipole("MUR", cells=(20, 20, 24), nrts=700, end=1e-12, nf2ff_freqs=nf2ff_freqs)
    if td is not None:
        F.nf2ff_td = td
    path = scenes.tmp_sim_path(tag)
    F.Run(path, cleanup=True)
    scenes.use_cuda_engine()
    return F, nf, port, path


THETA, PHI = np.arange(0.0, 181.0, 15.0), np.array([0.0, 90.0])


def _check_td_equals_running(engine):
    f_other = 2.4e9
    Fa, nfa, _, pa = _run(engine, [3e9, f_other], f"any_a_{engine}")          # f_other by the running DFT
    Fb, nfb, _, pb = _run(engine, None, f"any_b_{engine}")                    # only f0 registered: f_other from the stored samples
    assert Fb.sim.td_store and Fb.sim.td_bytes > 0
    ra = nfa.CalcNF2FF(pa, f_other, THETA, PHI, center=[0, 0, 0])
    rb = nfb.CalcNF2FF(pb, f_other, THETA, PHI, center=[0, 0, 0])
    assert abs(10 * np.log10(ra.Dmax[0]) - 10 * np.log10(rb.Dmax[0])) <= 0.01
    ea, eb = ra.E_norm[0], rb.E_norm[0]
    sel = ea > ea.max() * 1e-2
    assert np.abs(20 * np.log10(eb[sel] / ea[sel])).max() <= 0.01
    # the registered frequency still comes from the running DFT, and equals the transform of the stored samples
    sp_td = Fb.sim.nf2ff_spectra([3e9])
    for a_run, a_td in zip(Fb.results["nf2ff"]["acc"], sp_td):
        assert np.abs(a_run[:, 0] - a_td[:, 0]).max() <= 2e-5 * np.abs(a_run).max()
    # the per-phi loop of the reference reuses the cached sources (one entry per (frequency, centre))
    for ph in (0.0, 45.0, 90.0):
        nfb.CalcNF2FF(pb, f_other, THETA, np.array([ph]), center=[0, 0, 0])
    assert len(Fb.results["nf2ff"]["sources"]) == 1
    if engine == "oracle":          # host path: the transformed spectra are cached too (the CUDA path keeps them on the device)
        assert len(Fb.results["nf2ff"]["extra"]) == 1
    return rb


def test_any_frequency_from_stored_samples_oracle_engine():
    _check_td_equals_running("oracle")


def test_unregistered_frequency_without_store_is_an_error():
    F, nf, _, path = _run("oracle", None, "any_off", td=False)
    assert not F.sim.td_store
    with pytest.raises(ValueError):
        nf.CalcNF2FF(path, 7.77e9, np.array([0.0]), np.array([0.0]))
    nf.CalcNF2FF(path, 3e9, np.array([0.0]), np.array([0.0]))       # the registered one still works


@pytest.mark.gpu
def test_any_frequency_from_stored_samples_cuda_vs_oracle():
    rc = _check_td_equals_running("cuda")
    _, nfo, _, po = _run("oracle", None, "any_o2")
    ro = nfo.CalcNF2FF(po, 2.4e9, THETA, PHI, center=[0, 0, 0])
    assert abs(10 * np.log10(ro.Dmax[0]) - 10 * np.log10(rc.Dmax[0])) <= 0.1        # north-star gain tolerance
    eo, ec = ro.E_norm[0], rc.E_norm[0]
    sel = eo > eo.max() * 10 ** (-30 / 20)
    assert np.abs(20 * np.log10(ec[sel] / eo[sel])).max() <= 0.1
