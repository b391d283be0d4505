"""Post-processing of a finished run: port spectra (CalcPort) and the NF2FF transform (CalcNF2FF).

Formulas: SURVEY.md App. A5/A6; call sites in the reference:
  port.CalcPort(sim_path, f); s11 = port.uf_ref / port.uf_inc      antenna_sim/solver_fdtd_openems_microstrip.py:408-413
  nf2ff.CalcNF2FF(sim_path, f_res, theta, phi, center=…)            antenna_sim/solver_fdtd_openems_microstrip_3d.py:225
The radiation integral over all directions runs on the GPU (K11, csrc/b200fdtd.cu farfield_kernel).
"""
from __future__ import annotations

import numpy as np

from .constants import C0, Z0


def dft_time2freq(t, val, freq):
    """openEMS utilities.DFT_time2freq, signal_type 'pulse': 2*dt*sum(val*exp(-j 2 pi f t))"""
    t = np.asarray(t, np.float64); val = np.asarray(val, np.float64); freq = np.atleast_1d(np.asarray(freq, np.float64))
    if len(t) < 2:
        return np.zeros(len(freq), np.complex128)
    out = np.empty(len(freq), np.complex128)
    for n, f in enumerate(freq):
        out[n] = np.sum(val * np.exp(-2j * np.pi * f * t))
    return 2.0 * (t[1] - t[0]) * out


def port_spectrum(probe, freq, probe_freqs=None):
    """spectrum of one probe record (dict from Simulation.collect): device running DFT when `freq` is the
    registered grid, otherwise a host DFT of the stored time series (any frequency)."""
    freq = np.atleast_1d(np.asarray(freq, np.float64))
    if probe.get("dft") is not None and probe_freqs is not None and len(probe_freqs) == len(freq) \
            and np.allclose(probe_freqs, freq, rtol=1e-12, atol=0.0):
        dts = probe["t"][1] - probe["t"][0] if len(probe["t"]) > 1 else 0.0
        return 2.0 * dts * np.asarray(probe["dft"], np.complex128)
    return dft_time2freq(probe["t"], probe["val"], freq)


def face_spectra(nf, freq):
    """per-face [4][nb][na] complex spectra (Ea, Eb, Ha, Hb) at `freq`: the running-DFT accumulators if the frequency was
    registered, otherwise a DFT of the stored time-domain face samples (what openEMS's CalcNF2FF does with its HDF5 dumps;
    the reference passes the caller's frequency_hz: antenna_sim/solver_fdtd_openems_microstrip_3d.py:225)"""
    freqs = np.asarray(nf["freqs"], np.float64)
    hit = np.where(np.isclose(freqs, freq, rtol=1e-9, atol=0.0))[0]
    if len(hit):
        return [acc[:, int(hit[0])] for acc in nf["acc"]]
    key = float(freq)
    extra = nf.setdefault("extra", {})
    if key not in extra:
        fn = nf.get("spectra_fn")
        if fn is None:
            raise ValueError(f"NF2FF frequency {freq:g} Hz was not registered for the running DFT (available: {freqs.tolist()}) "
                             "and the time-domain face samples were not kept (memory budget B200FDTD_NF2FF_TD_GB); "
                             "pass frequency=[...] to CreateNF2FFBox")
        extra[key] = [a[:, 0] for a in fn([key])]
    return extra[key]


def surface_currents(nf, freq, center):
    """equivalent currents on the Huygens box at frequency `freq`.
    Returns pos [3][N] (relative to center, m), J [3][N], M [3][N] (already times dA), Prad."""
    pos, Jl, Ml = [], [], []
    prad = 0.0
    for F, acc, (xa, xb, wa, wb) in zip(nf["faces"], face_spectra(nf, freq), nf["weights"]):
        n = F["normal"]; a, b = (n + 1) % 3, (n + 2) % 3
        s = 1.0 if F["side"] == 1 else -1.0
        Ea, Eb, Ha, Hb = (acc[c] for c in range(4))                 # [nb][na]
        dA = wb[:, None] * wa[None, :]
        XA, XB = np.meshgrid(xa, xb)                                 # [nb][na]
        P = np.zeros((3,) + dA.shape)
        P[n] = F["coord"]; P[a] = XA; P[b] = XB
        J = np.zeros((3,) + dA.shape, np.complex128); M = np.zeros_like(J)
        J[a] = -s * Hb * dA; J[b] = s * Ha * dA                     # J = n x H
        M[a] = s * Eb * dA;  M[b] = -s * Ea * dA                    # M = -n x E
        prad += 0.5 * s * np.sum(np.real(Ea * np.conj(Hb) - Eb * np.conj(Ha)) * dA)
        pos.append(P.reshape(3, -1)); Jl.append(J.reshape(3, -1)); Ml.append(M.reshape(3, -1))
    c = np.asarray(center, np.float64).reshape(3)
    pos = np.concatenate(pos, 1) - c.reshape(3, 1)
    J, M = np.concatenate(Jl, 1), np.concatenate(Ml, 1)
    # image theory for box faces dropped on PEC / PMC walls (openEMS nf2ff 'mirror'): the image of an electric current in a
    # PEC wall keeps its normal and flips its tangential components, a magnetic current the other way round (PMC: swapped)
    for (n, wall, kind) in nf.get("mirrors", []):
        pm, Jm, Mm = pos.copy(), J.copy(), M.copy()
        pm[n] = 2.0 * (wall - c[n]) - pos[n]
        tang = [a for a in range(3) if a != n]
        if kind == 0:          # PEC
            Jm[tang] *= -1.0; Mm[n] *= -1.0
        else:                  # PMC
            Jm[n] *= -1.0; Mm[tang] *= -1.0
        pos, J, M = np.concatenate([pos, pm], 1), np.concatenate([J, Jm], 1), np.concatenate([M, Mm], 1)
    return pos, J, M, float(prad)


def device_sources(nf, freq, center):
    """equivalent currents formed ON THE DEVICE from this rank's device-resident face spectra (CUDA engine): the
    accumulators never travel to the host.  Same formulas as surface_currents.  Returns (FarfieldSources, Prad); on a
    z-slab run Prad is already summed over the ranks (collective)."""
    import torch
    from .engine import FarfieldSources
    D = nf["device"]
    dev = D["dev"]
    freqs = np.asarray(nf["freqs"], np.float64)
    hit = np.where(np.isclose(freqs, freq, rtol=1e-9, atol=0.0))[0]
    faces = D["faces"]
    if len(hit):
        accs = [F["acc"][:, int(hit[0])] for F in faces]             # [4][nb][na][2] float32 views
    else:
        if D.get("td_dft") is None:
            raise ValueError(f"NF2FF frequency {freq:g} Hz was not registered for the running DFT (available: {freqs.tolist()}) "
                             "and the time-domain face samples were not kept (memory budget B200FDTD_NF2FF_TD_GB); "
                             "pass frequency=[...] to CreateNF2FFBox")
        accs = [a[:, 0] for a in D["td_dft"](float(freq))] if faces else []
    scale = float(D["scale"])
    c = np.ascontiguousarray(np.asarray(center, np.float64).reshape(3))
    # one kernel per face (K11a, csrc: nf2ff_sources_kernel) through the C-ABI; the spectra are used where they lie
    from . import _lib
    import ctypes as C
    L = _lib.lib()
    npts = sum(int(a.shape[1]) * int(a.shape[2]) for a in accs)
    pos = torch.empty((3, npts), dtype=torch.float32, device=dev)
    J = torch.empty((3, npts, 2), dtype=torch.float32, device=dev)
    M = torch.empty((3, npts, 2), dtype=torch.float32, device=dev)
    arr = (_lib.Nf2ffSrcFace * max(1, len(faces)))()
    keep = []
    for q, (F, acc) in enumerate(zip(faces, accs)):
        nb_, na_ = int(acc.shape[1]), int(acc.shape[2])
        assert acc.stride(3) == 1 and acc.stride(2) == 2 and acc.stride(1) == 2 * na_, "face spectra must be [nb][na][2] rows"
        xs = [np.ascontiguousarray(F[k], np.float64) for k in ("xa", "xb", "wa", "wb")]
        keep.append(xs)
        arr[q].normal, arr[q].side, arr[q].na, arr[q].nb = int(F["normal"]), int(F["side"]), na_, nb_
        arr[q].coord = float(F["coord"])
        arr[q].acc = acc.data_ptr(); arr[q].comp_stride = int(acc.stride(0))
        arr[q].xa, arr[q].xb, arr[q].wa, arr[q].wb = (x.ctypes.data_as(_lib.c_d) for x in xs)
    prad_h = C.c_double(0.0)
    st = torch.cuda.current_stream(dev)
    _lib.check(L.b200fdtd_nf2ff_sources(dev.index, C.c_void_p(st.cuda_stream), len(faces), arr, scale, c.ctypes.data_as(_lib.c_d),
                                        pos.data_ptr(), J.data_ptr(), M.data_ptr(), C.byref(prad_h)))
    prad = torch.tensor([prad_h.value], dtype=torch.float64, device=dev)
    if D.get("world", 1) > 1:
        torch.distributed.all_reduce(prad, group=D.get("group"))
    src = FarfieldSources.from_device(pos, J, M)
    src.world, src.group = D.get("world", 1), D.get("group")
    return src, float(prad.item())


def far_field(nf, freq, theta_deg, phi_deg, center=(0, 0, 0), radius=1.0, farfield_fn=None, device=0):
    """E_theta/E_phi on the theta x phi grid (degrees) for one frequency of the box spectra."""
    # the reference calls CalcNF2FF once per phi with the same frequency and centre (73 calls): the equivalent currents are
    # formed once per (frequency, centre) and, on the CUDA path, stay on the device between the calls
    key = (float(freq), tuple(np.asarray(center, np.float64).reshape(3).tolist()))
    cache = nf.setdefault("sources", {})
    if key not in cache:
        if len(cache) >= 8:
            cache.clear()
        if farfield_fn is None and nf.get("device") is not None:
            src, prad = device_sources(nf, float(freq), center)
            cache[key] = dict(cur=(None, None, None, prad), dev=src)
        else:
            cache[key] = dict(cur=surface_currents(nf, float(freq), center), dev=None)
    ent = cache[key]
    pos, J, M, prad = ent["cur"]
    th = np.deg2rad(np.atleast_1d(np.asarray(theta_deg, np.float64)))
    ph = np.deg2rad(np.atleast_1d(np.asarray(phi_deg, np.float64)))
    TH, PH = np.meshgrid(th, ph, indexing="ij")
    k = 2.0 * np.pi * float(freq) / C0
    if farfield_fn is None:
        from .engine import farfield as farfield_fn, FarfieldSources     # CUDA kernel K11
        if ent["dev"] is None:
            ent["dev"] = FarfieldSources(pos, J, M, device)
        Nt, Np_, Lt, Lp = farfield_fn(None, None, None, k, TH.ravel(), PH.ravel(), device=device, sources=ent["dev"])
    else:
        Nt, Np_, Lt, Lp = farfield_fn(pos, J, M, k, TH.ravel(), PH.ravel())
    fac = 1j * k * np.exp(-1j * k * radius) / (4.0 * np.pi * radius)
    E_theta = (-fac * (Lp + Z0 * Nt)).reshape(TH.shape)
    E_phi = (fac * (Lt - Z0 * Np_)).reshape(TH.shape)
    P_rad = (np.abs(E_theta) ** 2 + np.abs(E_phi) ** 2) / (2.0 * Z0)
    Dmax = 4.0 * np.pi * radius ** 2 * P_rad.max() / prad if prad > 0 else float("nan")
    return dict(E_theta=E_theta, E_phi=E_phi, P_rad=P_rad, Prad=prad, Dmax=Dmax)
