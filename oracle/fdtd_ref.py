"""ctypes front-end of the CPU oracle (oracle/fdtd_ref.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.  PARITY UNPINNED (see fdtd_ref.c header).

The oracle consumes exactly the engine-level inputs of include/b200fdtd.h (same layout, same
linear indices), so one set of numpy arrays drives both the oracle and the CUDA engine.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_f = C.POINTER(C.c_float)
c_d = C.POINTER(C.c_double)
c_i64 = C.POINTER(C.c_int64)
c_i32 = C.POINTER(C.c_int32)


class _PmlBox(C.Structure):
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("z0", C.c_int32),
                ("bx", C.c_int32), ("by", C.c_int32), ("bz", C.c_int32),
                ("flux_v", c_f), ("flux_i", c_f), ("vv", c_f), ("vvfo", c_f), ("vvfn", c_f),
                ("ii", c_f), ("iifo", c_f), ("iifn", c_f)]


class _Face(C.Structure):
    _fields_ = [("normal", C.c_int32), ("plane", C.c_int32), ("a0", C.c_int32), ("a1", C.c_int32),
                ("b0", C.c_int32), ("b1", C.c_int32), ("acc", c_d), ("td", c_f), ("td_max", C.c_int32)]


class _Engine(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("px", C.c_int32),
                ("volt", c_f), ("curr", c_f), ("vv", c_f), ("vi", c_f), ("ii", c_f), ("iv", c_f),
                ("n_exc", C.c_int64), ("exc_idx", c_i64), ("exc_amp", c_f), ("exc_delay", c_i32),
                ("exc_sig", c_f), ("exc_siglen", C.c_int32),
                ("n_mur", C.c_int64), ("mur_dst", c_i64), ("mur_src", c_i64), ("mur_coeff", c_f), ("mur_tmp", c_f),
                ("n_pml", C.c_int32), ("pml", C.POINTER(_PmlBox)),
                ("n_probes", C.c_int32), ("pr_kind", c_i32), ("pr_off", c_i64), ("pr_idx", c_i64), ("pr_w", c_f),
                ("interval", C.c_int32), ("max_samples", C.c_int32), ("pr_series", c_d), ("pr_nfreq", C.c_int32),
                ("pr_freqs", c_d), ("pr_dft", c_d), ("dt", C.c_double),
                ("n_faces", C.c_int32), ("faces", C.POINTER(_Face)), ("nf_nfreq", C.c_int32), ("nf_freqs", c_d),
                ("inv_len", c_f * 3), ("inv_dual", c_f * 3),
                ("ts", C.c_int64), ("threads", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile oracle/fdtd_ref.c -> oracle/libfdtd_ref.so (gcc; see Makefile)."""
    so = os.path.join(_HERE, "libfdtd_ref.so")
    src = os.path.join(_HERE, "fdtd_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfdtd_ref.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libfdtd_ref.so")
        src = os.path.join(_HERE, "fdtd_ref.c")
        if not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(so) < os.path.getmtime(src)):
            build()
        L = C.CDLL(so)
        for name in ("ref_update_e", "ref_update_h", "ref_excite", "ref_step", "ref_probes", "ref_nf2ff"):
            getattr(L, name).argtypes = [C.POINTER(_Engine)]
            getattr(L, name).restype = None
        L.ref_mur.argtypes = [C.POINTER(_Engine), C.c_int]
        L.ref_pml.argtypes = [C.POINTER(_Engine), C.c_int, C.c_int]
        L.ref_half_step.argtypes = [C.POINTER(_Engine), C.c_int]
        L.ref_half_step_part.argtypes = [C.POINTER(_Engine), C.c_int, C.c_int]
        L.ref_run.argtypes = [C.POINTER(_Engine), C.c_int64]
        L.ref_energy.argtypes = [C.POINTER(_Engine)]
        L.ref_energy.restype = C.c_double
        L.ref_farfield.argtypes = [C.c_int64, c_d, c_d, c_d, C.c_double, C.c_int, c_d, c_d, c_d]
        L.ref_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t)


def _host(a, dt=np.float32):
    """numpy view/copy of a numpy array or a (possibly CUDA) torch tensor"""
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dt)


def lin(nz, ny, px, c, k, j, i):
    """linear index of (component c, local plane k, row j, column i) — include/b200fdtd.h"""
    return ((c * (nz + 2) + (k + 1)) * ny + j) * px + i


class RefEngine:
    """numpy-backed oracle engine with the same inputs as b200fdtd.Engine."""

    def __init__(self, nx, ny, nz, px=None, threads=1):
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.px = int(px) if px else (self.nx + 3) // 4 * 4
        shape = (3, self.nz + 2, self.ny, self.px)
        self.shape = shape
        self.volt = np.zeros(shape, np.float32)
        self.curr = np.zeros(shape, np.float32)
        self.vv = np.zeros(shape, np.float32)
        self.vi = np.zeros(shape, np.float32)
        self.ii = np.zeros(shape, np.float32)
        self.iv = np.zeros(shape, np.float32)
        self._keep = {}
        self.series = np.zeros((0, 0)); self.probe_dft = np.zeros((0, 0, 2)); self.face_acc = []
        self.e = _Engine()
        self.e.nx, self.e.ny, self.e.nz, self.e.px = self.nx, self.ny, self.nz, self.px
        self.e.threads = int(threads)
        self._bind()

    def _bind(self):
        e = self.e
        e.volt, e.curr = _p(self.volt, c_f), _p(self.curr, c_f)
        e.vv, e.vi, e.ii, e.iv = (_p(a, c_f) for a in (self.vv, self.vi, self.ii, self.iv))

    def set_coeffs(self, vv, vi, ii, iv):
        for dst, src in ((self.vv, vv), (self.vi, vi), (self.ii, ii), (self.iv, iv)):
            dst[...] = _host(src).reshape(self.shape)

    def set_excitation(self, idx, amp, delay, signal):
        k = self._keep
        k["exc_idx"] = np.ascontiguousarray(idx, np.int64)
        k["exc_amp"] = np.ascontiguousarray(amp, np.float32)
        k["exc_delay"] = np.ascontiguousarray(delay, np.int32)
        k["exc_sig"] = np.ascontiguousarray(signal, np.float32)
        e = self.e
        e.n_exc = len(k["exc_idx"]); e.exc_idx = _p(k["exc_idx"], c_i64); e.exc_amp = _p(k["exc_amp"], c_f)
        e.exc_delay = _p(k["exc_delay"], c_i32); e.exc_sig = _p(k["exc_sig"], c_f); e.exc_siglen = len(k["exc_sig"])

    def set_mur(self, dst, src, coeff):
        k = self._keep
        k["mur_dst"] = np.ascontiguousarray(dst, np.int64)
        k["mur_src"] = np.ascontiguousarray(src, np.int64)
        k["mur_coeff"] = np.ascontiguousarray(coeff, np.float32)
        k["mur_tmp"] = np.zeros(len(k["mur_dst"]), np.float32)
        e = self.e
        e.n_mur = len(k["mur_dst"]); e.mur_dst = _p(k["mur_dst"], c_i64); e.mur_src = _p(k["mur_src"], c_i64)
        e.mur_coeff = _p(k["mur_coeff"], c_f); e.mur_tmp = _p(k["mur_tmp"], c_f)

    def set_pml(self, boxes):
        """boxes: list of dicts {x0,y0,z0,bx,by,bz, vv,vvfo,vvfn,ii,iifo,iifn: arrays [3][bz][by][bx]}"""
        arr = (_PmlBox * len(boxes))()
        keep = []
        for b, B in enumerate(boxes):
            shp = (3, B["bz"], B["by"], B["bx"])
            d = {n: np.ascontiguousarray(_host(B[n]).reshape(shp)) for n in ("vv", "vvfo", "vvfn", "ii", "iifo", "iifn")}
            d["flux_v"] = np.zeros(shp, np.float32); d["flux_i"] = np.zeros(shp, np.float32)
            keep.append(d)
            for n in ("x0", "y0", "z0", "bx", "by", "bz"):
                setattr(arr[b], n, int(B[n]))
            for n, a in d.items():
                setattr(arr[b], n, _p(a, c_f))
        self._keep["pml"] = (arr, keep)
        self.pml_arrays = keep
        self.e.n_pml = len(boxes); self.e.pml = arr

    def set_probes(self, kind, offset, idx, weight, interval, max_samples, freqs, dt):
        k = self._keep
        k["pr_kind"] = np.ascontiguousarray(kind, np.int32); k["pr_off"] = np.ascontiguousarray(offset, np.int64)
        k["pr_idx"] = np.ascontiguousarray(idx, np.int64); k["pr_w"] = np.ascontiguousarray(weight, np.float32)
        k["pr_freqs"] = np.ascontiguousarray(freqs, np.float64)
        n = len(k["pr_kind"])
        self.series = np.zeros((n, max_samples), np.float64)
        self.probe_dft = np.zeros((n, len(k["pr_freqs"]), 2), np.float64)
        e = self.e
        e.n_probes = n; e.pr_kind = _p(k["pr_kind"], c_i32); e.pr_off = _p(k["pr_off"], c_i64)
        e.pr_idx = _p(k["pr_idx"], c_i64); e.pr_w = _p(k["pr_w"], c_f)
        e.interval = int(interval); e.max_samples = int(max_samples); e.pr_series = _p(self.series, c_d)
        e.pr_nfreq = len(k["pr_freqs"]); e.pr_freqs = _p(k["pr_freqs"], c_d); e.pr_dft = _p(self.probe_dft, c_d)
        e.dt = float(dt)

    def set_nf2ff(self, faces, freqs, interval, dt, inv_len, inv_dual):
        """faces: list of dicts {normal, plane, a0, a1, b0, b1}; inv_len/inv_dual: 3 float arrays (z: nz+2 entries)"""
        k = self._keep
        k["nf_freqs"] = np.ascontiguousarray(freqs, np.float64)
        nf = len(k["nf_freqs"])
        arr = (_Face * len(faces))()
        self.face_acc = []
        for q, F in enumerate(faces):
            na, nb = F["a1"] - F["a0"] + 1, F["b1"] - F["b0"] + 1
            acc = np.zeros((4, nf, nb, na, 2), np.float64)
            self.face_acc.append(acc)
            for n in ("normal", "plane", "a0", "a1", "b0", "b1"):
                setattr(arr[q], n, int(F[n]))
            arr[q].acc = _p(acc, c_d)
        k["faces"] = arr
        k["il"] = [np.ascontiguousarray(a, np.float32) for a in inv_len]
        k["idl"] = [np.ascontiguousarray(a, np.float32) for a in inv_dual]
        e = self.e
        e.n_faces = len(faces); e.faces = arr; e.nf_nfreq = nf; e.nf_freqs = _p(k["nf_freqs"], c_d)
        for a in range(3):
            e.inv_len[a] = _p(k["il"][a], c_f); e.inv_dual[a] = _p(k["idl"][a], c_f)
        e.interval = int(interval); e.dt = float(dt)

    def set_nf2ff_td(self, max_samples):
        """time-domain store of the face samples (mirrors b200fdtd_set_nf2ff_td)"""
        arr = self._keep["faces"]
        self.face_td = [np.zeros((int(max_samples), 4) + acc.shape[2:4], np.float32) for acc in self.face_acc]
        self.td_max = int(max_samples)
        for q, t in enumerate(self.face_td):
            arr[q].td = _p(t, c_f); arr[q].td_max = int(max_samples)
        return sum(t.nbytes for t in self.face_td)

    def nf2ff_td_dft(self, freqs, nsamples):
        """double-precision DFT of the stored samples (mirrors b200fdtd_nf2ff_td_dft): list of [4][nfreq][nb][na][2]"""
        freqs = np.atleast_1d(np.asarray(freqs, np.float64))
        ns = int(min(nsamples, self.td_max))
        iv, dt = int(self.e.interval), float(self.e.dt)
        s = np.arange(1, ns + 1, dtype=np.float64) * iv
        out = []
        for t in self.face_td:
            o = np.zeros((4, len(freqs)) + t.shape[2:4] + (2,), np.float64)
            x = t[:ns].astype(np.float64)
            for fi, f in enumerate(freqs):
                for c in range(4):
                    tm = (s + (0.5 if c >= 2 else 0.0)) * dt
                    ph = f * tm; ph -= np.floor(ph)
                    w = np.exp(-2j * np.pi * ph)
                    z = np.tensordot(w, x[:, c], axes=(0, 0))
                    o[c, fi, ..., 0] = z.real; o[c, fi, ..., 1] = z.imag
            out.append(o)
        return out

    def reset_state(self):
        for a in [self.volt, self.curr, self.series, self.probe_dft] + list(self.face_acc) + list(getattr(self, "face_td", [])):
            a[...] = 0
        if "pml" in self._keep:
            for d in self._keep["pml"][1]:
                d["flux_v"][...] = 0; d["flux_i"][...] = 0
        if "mur_tmp" in self._keep:
            self._keep["mur_tmp"][...] = 0
        self.e.ts = 0

    # --- stepping ---
    def run(self, nsteps, use_graph=False):
        lib().ref_run(C.byref(self.e), int(nsteps))

    @property
    def num_samples(self):
        return int(self.e.ts // self.e.interval) if self.e.interval > 0 else 0

    def sync(self):
        pass

    def half_step(self, phase):
        lib().ref_half_step(C.byref(self.e), int(phase))

    def half_step_part(self, phase, part):
        self._load_current()
        lib().ref_half_step_part(C.byref(self.e), int(phase), int(part))
        self._store_current()

    def half_step_raw(self, phase):
        if phase == 3:
            # sampling between two fused steps: H from the current copy, E from the copy that is NOT current (it still holds
            # the voltages of the step being sampled, with the ghost planes the caller has exchanged into it)
            if getattr(self, "volt2", None) is None:
                raise RuntimeError("pipelined sampling needs the second field copy")
            if self._ccur:
                self.curr[...] = self.curr2
            keep = self.volt.copy()
            if self._vcur == 0:
                self.volt[...] = self.volt2          # old copy = copy 1
            lib().ref_half_step(C.byref(self.e), 2)
            self.volt[...] = keep
            if self._ccur:
                self.curr[:, 1:self.nz + 1] = np.nan
            return
        self._load_current()
        self.half_step(phase)
        self._store_current()

    # --- protocol emulation of the CUDA engine's fused steps on a z-slab rank (b200fdtd_fused_step_part) ---------------
    # The C oracle always updates volt/curr in place.  To let the CPU gloo tests exercise the HOST side of the protocol
    # (simulation.py:_fused_step: which field copy is current, which copy is sent / received when), this keeps a second
    # copy on the python side and moves the state between the copies exactly where the CUDA launches read and write them.
    # A copy that does not hold the state is filled with NaN, so reading or sending the wrong copy cannot go unnoticed.
    emulates_fused_steps = True

    def bind_alt_fields(self):
        self.volt2 = np.zeros_like(self.volt); self.curr2 = np.zeros_like(self.curr)
        self._vcur = self._ccur = 0

    def current_copy(self):
        return self._vcur, self._ccur

    def reset_current_copy(self):
        self._vcur = self._ccur = 0

    def _load_current(self):
        """engine arrays <- the copies that hold the state, ghost planes included (halos land in the current copy)"""
        if getattr(self, "volt2", None) is None:
            return
        if self._vcur:
            self.volt[...] = self.volt2
        if self._ccur:
            self.curr[...] = self.curr2

    def _store_current(self):
        """owned planes back into the copies that hold the state; the engine arrays' owned planes are poisoned then"""
        if getattr(self, "volt2", None) is None:
            return
        o = slice(1, self.nz + 1)
        if self._vcur:
            self.volt2[:, o] = self.volt[:, o]; self.volt[:, o] = np.nan
        if self._ccur:
            self.curr2[:, o] = self.curr[:, o]; self.curr[:, o] = np.nan

    def fused_step_part(self, part):
        nz = self.nz
        o = slice(1, nz + 1)
        if part == 0:                                   # H of planes 0 .. nz-2 (CUDA engine: plane 0 + slabs here, interior in part 4/2)
            self._load_current()
            lib().ref_half_step_part(C.byref(self.e), 1, 0)
        elif part == 4:
            pass
        elif part == 1:                                 # H of the top plane: needs the upper ghost E of the CURRENT E copy
            if self._vcur:
                self.volt[:, nz + 1] = self.volt2[:, nz + 1]
            lib().ref_half_step_part(C.byref(self.e), 1, 1)
            if self._ccur == 0:                         # the new H belongs in copy 1: the caller sends it from there
                self.curr2[:, o] = self.curr[:, o]; self.curr[:, o] = np.nan
            else:                                       # it belongs in copy 0 (the engine array); copy 1 is stale from here on
                self.curr2[:, o] = np.nan
        elif part == 2:
            self._ccur ^= 1
        elif part == 3:                                 # E: needs the new H with its freshly received lower ghost plane
            if self._ccur:
                self.curr[...] = self.curr2             # owned planes parked in part 1 + ghost plane 0 received since
            if self._vcur:
                self.volt[...] = self.volt2             # the current E (copy 1) into the engine array, ghost planes included
            old = self.volt[:, o].copy()                # the fused launch only reads the old E: its copy keeps holding it
            lib().ref_half_step_part(C.byref(self.e), 0, 0)
            lib().ref_half_step_part(C.byref(self.e), 0, 1)
            if self._ccur:
                self.curr[:, o] = np.nan                # H was only read: copy 1 still holds it
            if self._vcur == 0:                         # the new E belongs in copy 1, copy 0 keeps the old one
                self.volt2[:, o] = self.volt[:, o]; self.volt[:, o] = old
            else:                                       # it belongs in copy 0 (the engine array), copy 1 keeps the old one
                self.volt2[:, o] = old
            self._vcur ^= 1
        else:
            raise ValueError("part must be 0..4")

    def update_only(self, which):
        (lib().ref_update_e if which == 0 else lib().ref_update_h)(C.byref(self.e))

    def energy(self):
        self._load_current()
        e = float(lib().ref_energy(C.byref(self.e)))
        self._store_current()
        return e

    @property
    def ts(self):
        return int(self.e.ts)


def farfield(pos, J, M, k, theta, phi):
    """pos [3][n] f64, J/M [3][n] complex128, theta/phi radians -> (N_theta, N_phi, L_theta, L_phi) complex [ndir]"""
    pos = np.ascontiguousarray(pos, np.float64)
    n = pos.shape[1]
    Jc = np.ascontiguousarray(np.stack([np.real(J), np.imag(J)], -1), np.float64)
    Mc = np.ascontiguousarray(np.stack([np.real(M), np.imag(M)], -1), np.float64)
    th = np.ascontiguousarray(theta, np.float64); ph = np.ascontiguousarray(phi, np.float64)
    out = np.zeros((len(th), 4, 2), np.float64)
    lib().ref_farfield(n, _p(pos, c_d), _p(Jc, c_d), _p(Mc, c_d), float(k), len(th), _p(th, c_d), _p(ph, c_d), _p(out, c_d))
    oc = out[..., 0] + 1j * out[..., 1]
    return oc[:, 0], oc[:, 1], oc[:, 2], oc[:, 3]
