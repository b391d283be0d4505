"""Engine: host-side owner of one z-slab of the FDTD state on one B200.

PyTorch provides device memory and the stream; every arithmetic step is a hand-written
sm_100a kernel in csrc/b200fdtd.cu reached through the C-ABI (include/b200fdtd.h).
This object is what `openEMS.openEMS.Run` drives (the reference's FDTD.Run call,
antenna_sim/solver_fdtd_openems_microstrip_3d.py:214).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, c_f, c_d, c_i32, c_i64


def _np(a, dt):
    return np.ascontiguousarray(a, dt)


def _ptr(a, t):
    return a.ctypes.data_as(t)


def round_up(n, m):
    return (n + m - 1) // m * m


def lin(nz, ny, px, c, k, j, i):
    """linear index of (component c, local plane k, row j, column i) — include/b200fdtd.h"""
    return ((c * (nz + 2) + (k + 1)) * ny + j) * px + i


class Engine:
    """One z-slab on one GPU.  Arrays are torch float32 CUDA tensors of shape [3, nz+2, ny, px]."""

    def __init__(self, nx, ny, nz, px=None, device=0, stream=None):
        if not torch.cuda.is_available():
            raise _lib.B200FDTDError("no CUDA device: the B200 engine has no CPU fallback")
        self.L = _lib.lib()
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.px = int(px) if px else round_up(self.nx, 32)
        self.device = torch.device("cuda", int(device))
        self.shape = (3, self.nz + 2, self.ny, self.px)
        torch.cuda.set_device(self.device)
        # a dedicated stream (graph capture needs a non-default stream); joined with torch's current
        # stream before and after every stepping call so tensor reads/writes stay ordered
        self.stream = stream if stream is not None else torch.cuda.Stream(self.device)
        self.h = C.c_void_p()
        check(self.L.b200fdtd_create(C.byref(self.h), self.device.index, self.nx, self.ny, self.nz, self.px,
                                      C.c_void_p(self.stream.cuda_stream)))
        self.volt = torch.zeros(self.shape, dtype=torch.float32, device=self.device)
        self.curr = torch.zeros(self.shape, dtype=torch.float32, device=self.device)
        check(self.L.b200fdtd_bind_fields(self.h, self.volt.data_ptr(), self.curr.data_ptr()))
        self.vv = self.vi = self.ii = self.iv = None
        self._keep = {}
        self.series = None
        self.probe_dft = None
        self.face_acc = []

    # ---- lifetime ----
    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.L.b200fdtd_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- inputs ----
    def _dev(self, a):
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.device, dtype=torch.float32)
        else:
            t = torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(self.device)
        return t.contiguous()

    def set_coeffs(self, vv, vi, ii, iv):
        self.vv, self.vi, self.ii, self.iv = (self._dev(a).reshape(self.shape) for a in (vv, vi, ii, iv))
        check(self.L.b200fdtd_bind_coeffs(self.h, self.vv.data_ptr(), self.vi.data_ptr(),
                                          self.ii.data_ptr(), self.iv.data_ptr()))

    def set_row_compression(self, which, xvecs, meta):
        """which 0 = E pass (vv, vi), 1 = H pass (ii, iv); xvecs float32 [nvec, px]; meta uint8 [nz+2, ny, 32].
        Returns (row-slots compressed, row-slots demoted by the device-side verification)."""
        if xvecs is None or len(xvecs) == 0:
            check(self.L.b200fdtd_set_row_compression(self.h, int(which), 0, None, None, None, None))
            return 0, 0
        xv = self._dev(xvecs).reshape(-1, self.px).contiguous()
        mt = meta if isinstance(meta, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(meta, np.uint8))
        mt = mt.to(self.device).contiguous()
        assert mt.dtype == torch.uint8 and mt.numel() == (self.nz + 2) * self.ny * 32
        self._keep[f"cmp{which}"] = (xv, mt)
        nc, nd = C.c_int64(), C.c_int64()
        self._pre()          # the verification kernel reads the coefficient arrays: order it after whatever filled them
        check(self.L.b200fdtd_set_row_compression(self.h, int(which), xv.shape[0], xv.data_ptr(), mt.data_ptr(),
                                                  C.byref(nc), C.byref(nd)))
        return nc.value, nd.value

    def expand_rows(self, which, xvecs, meta):
        """fill the bound full arrays of one pass from compression tables on the device (rows streamed in full become 0)"""
        self._pre()
        check(self.L.b200fdtd_expand_rows(self.h, int(which), int(xvecs.shape[0]), xvecs.data_ptr(), meta.data_ptr()))
        self._post()

    def set_tuning(self, kz=16, ty=4, variant=0):
        check(self.L.b200fdtd_set_tuning(self.h, int(kz), int(ty), int(variant)))

    def set_he_tuning(self, rows=0, planes=0, de=0):
        """tile of the fused H->E launch: rows per CTA in {3, 7, 15}, planes marched per CTA, E planes landing ahead of the two
        in use in {1, 2} (0 = keep / automatic)"""
        check(self.L.b200fdtd_set_he_tuning(self.h, int(rows), int(planes) | (int(de) << 24)))

    @property
    def he_active(self):
        """True if the last graph run used the fused H->E launches"""
        v = C.c_int()
        check(self.L.b200fdtd_he_info(self.h, C.byref(v)))
        return bool(v.value)

    def set_excitation(self, idx, amp, delay, signal):
        idx, amp, delay, signal = _np(idx, np.int64), _np(amp, np.float32), _np(delay, np.int32), _np(signal, np.float32)
        check(self.L.b200fdtd_set_excitation(self.h, len(idx), _ptr(idx, c_i64), _ptr(amp, c_f), _ptr(delay, c_i32),
                                             _ptr(signal, c_f), len(signal)))

    def set_mur(self, dst, src, coeff):
        dst, src, coeff = _np(dst, np.int64), _np(src, np.int64), _np(coeff, np.float32)
        check(self.L.b200fdtd_set_mur(self.h, len(dst), _ptr(dst, c_i64), _ptr(src, c_i64), _ptr(coeff, c_f)))

    def set_pml(self, boxes):
        """boxes: list of dicts {x0,y0,z0,bx,by,bz, vv,vvfo,vvfn,ii,iifo,iifn: arrays [3][bz][by][bx]}"""
        arr = (_lib.PmlBox * max(1, len(boxes)))()
        keep = []
        for b, B in enumerate(boxes):
            shp = (3, int(B["bz"]), int(B["by"]), int(B["bx"]))
            d = {n: self._dev(B[n]).reshape(shp).contiguous() for n in ("vv", "vvfo", "vvfn", "ii", "iifo", "iifn")}
            d["flux_v"] = torch.zeros(shp, dtype=torch.float32, device=self.device)
            d["flux_i"] = torch.zeros(shp, dtype=torch.float32, device=self.device)
            keep.append(d)
            for n in ("x0", "y0", "z0", "bx", "by", "bz"):
                setattr(arr[b], n, int(B[n]))
            for n, t in d.items():
                setattr(arr[b], n, t.data_ptr())
            for kind in ("v", "i"):                       # optional row compression of the slab coefficients
                cmp = B.get("cmp_" + kind)
                if cmp is not None and cmp[0] is not None:
                    xv = self._dev(cmp[0]).reshape(-1, shp[3]).contiguous()
                    mt = cmp[1] if isinstance(cmp[1], torch.Tensor) else torch.from_numpy(np.ascontiguousarray(cmp[1], np.uint8))
                    mt = mt.to(self.device).contiguous()
                    assert mt.dtype == torch.uint8 and mt.numel() == shp[1] * shp[2] * 48
                    d["xv_" + kind], d["meta_" + kind] = xv, mt
                    setattr(arr[b], "xvecs_" + kind, xv.data_ptr()); setattr(arr[b], "meta_" + kind, mt.data_ptr())
                    setattr(arr[b], "nvec_" + kind, int(xv.shape[0]))
        self._keep["pml"] = keep
        self.pml_arrays = keep
        self._pre()              # the verification kernels read the slab arrays: order them after whatever filled them
        check(self.L.b200fdtd_set_pml(self.h, len(boxes), arr))
        a, bb = C.c_int64(), C.c_int64()
        check(self.L.b200fdtd_pml_compression_info(self.h, C.byref(a), C.byref(bb)))
        self.pml_compression = (a.value, bb.value)

    def set_probes(self, kind, offset, idx, weight, interval, max_samples, freqs, dt):
        kind, offset = _np(kind, np.int32), _np(offset, np.int64)
        idx, weight, freqs = _np(idx, np.int64), _np(weight, np.float32), _np(freqs, np.float64)
        n = len(kind)
        self.series = torch.zeros((n, int(max_samples)), dtype=torch.float32, device=self.device)
        self.probe_dft = torch.zeros((n, max(1, len(freqs)), 2), dtype=torch.float32, device=self.device)
        self.probe_freqs = freqs
        self.interval = int(interval)
        check(self.L.b200fdtd_set_probes(self.h, n, _ptr(kind, c_i32), _ptr(offset, c_i64), _ptr(idx, c_i64),
                                         _ptr(weight, c_f), int(interval), int(max_samples), self.series.data_ptr(),
                                         len(freqs), _ptr(freqs, c_d), self.probe_dft.data_ptr(), float(dt)))

    def set_nf2ff(self, faces, freqs, interval, dt, inv_len, inv_dual):
        """faces: list of dicts {normal, plane, a0, a1, b0, b1}; inv_len/inv_dual: 3 float arrays (z: nz+2 entries)"""
        freqs = _np(freqs, np.float64)
        arr = (_lib.Nf2ffFace * max(1, len(faces)))()
        self.face_acc = []
        for q, F in enumerate(faces):
            na, nb = F["a1"] - F["a0"] + 1, F["b1"] - F["b0"] + 1
            acc = torch.zeros((4, len(freqs), nb, na, 2), dtype=torch.float32, device=self.device)
            self.face_acc.append(acc)
            for n in ("normal", "plane", "a0", "a1", "b0", "b1"):
                setattr(arr[q], n, int(F[n]))
            arr[q].acc = acc.data_ptr()
        il = [_np(a, np.float32) for a in inv_len]
        idl = [_np(a, np.float32) for a in inv_dual]
        self.nf_freqs = freqs
        self.interval = int(interval)
        check(self.L.b200fdtd_set_nf2ff(self.h, len(faces), arr, len(freqs), _ptr(freqs, c_d), int(interval), float(dt),
                                        _ptr(il[0], c_f), _ptr(il[1], c_f), _ptr(il[2], c_f),
                                        _ptr(idl[0], c_f), _ptr(idl[1], c_f), _ptr(idl[2], c_f)))

    def set_nf2ff_td(self, max_samples):
        """keep the face samples themselves in HBM (openEMS's nf2ff_E/H_n.h5 dumps) so the far field can be evaluated at any
        frequency after the run; returns the bytes allocated"""
        self.face_td = [torch.zeros((int(max_samples), 4) + tuple(acc.shape[2:4]), dtype=torch.float32, device=self.device)
                        for acc in self.face_acc]
        self.td_max = int(max_samples)
        arr = (C.c_void_p * max(1, len(self.face_td)))(*[t.data_ptr() for t in self.face_td])
        self._pre()
        check(self.L.b200fdtd_set_nf2ff_td(self.h, len(self.face_td), arr, int(max_samples)))
        return sum(t.numel() * 4 for t in self.face_td)

    def nf2ff_td_dft(self, freqs, nsamples):
        """DFT of the stored face samples at `freqs`: list of float32 device tensors [4][nfreq][nb][na][2] (accumulator layout)"""
        freqs = _np(np.atleast_1d(freqs), np.float64)
        out = []
        self._pre()
        for q, acc in enumerate(self.face_acc):
            o = torch.empty((4, len(freqs)) + tuple(acc.shape[2:4]) + (2,), dtype=torch.float32, device=self.device)
            check(self.L.b200fdtd_nf2ff_td_dft(self.h, q, len(freqs), _ptr(freqs, c_d), int(min(nsamples, self.td_max)), o.data_ptr()))
            out.append(o)
        self._post()
        return out

    def reset_state(self):
        """fields, PML flux, probe series / DFT and NF2FF accumulators back to zero, step counter to 0 (a new run of the same scene)"""
        self._pre()
        with torch.cuda.stream(self.stream):
            for t in (self.volt, self.curr, getattr(self, "volt2", None), getattr(self, "curr2", None), self.series, self.probe_dft):
                if t is not None:
                    t.zero_()
            for d in self._keep.get("pml", []):
                d["flux_v"].zero_(); d["flux_i"].zero_()
            for t in list(self.face_acc) + list(getattr(self, "face_td", [])):
                t.zero_()
        self.set_timestep(0)
        self._post()

    # ---- stepping ----
    def _pre(self):
        self.stream.wait_stream(torch.cuda.current_stream(self.device))

    def _post(self):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def run(self, nsteps, use_graph=True):
        self._pre()
        check(self.L.b200fdtd_run(self.h, int(nsteps), 1 if use_graph else 0))
        self._post()

    def half_step(self, phase):
        self._pre()
        check(self.L.b200fdtd_half_step(self.h, int(phase)))
        self._post()

    def half_step_part(self, phase, part):
        """no stream joins: the caller orders work on self.stream itself (z-slab overlap, simulation._step_multi)"""
        check(self.L.b200fdtd_half_step_part(self.h, int(phase), int(part)))

    # ---- fused H->E steps on a z-slab rank (b200fdtd_fused_step_part) ----
    def bind_alt_fields(self):
        """allocate and bind the second copy of the fields (caller-owned variant: the halo exchange needs the tensors)"""
        self.volt2 = torch.zeros(self.shape, dtype=torch.float32, device=self.device)
        self.curr2 = torch.zeros(self.shape, dtype=torch.float32, device=self.device)
        check(self.L.b200fdtd_bind_alt_fields(self.h, self.volt2.data_ptr(), self.curr2.data_ptr()))

    def fused_step_part(self, part):
        check(self.L.b200fdtd_fused_step_part(self.h, int(part)))

    def current_copy(self):
        a, b = C.c_int(), C.c_int()
        check(self.L.b200fdtd_current_copy(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def reset_current_copy(self):
        check(self.L.b200fdtd_reset_current_copy(self.h))

    def half_step_raw(self, phase):
        check(self.L.b200fdtd_half_step(self.h, int(phase)))

    def update_only(self, which, join=True):
        if join:
            self._pre()
        check(self.L.b200fdtd_update_only(self.h, int(which)))
        if join:
            self._post()

    def plan_info(self):
        """(plain_cells, fused_pml_cells, separate_pml_cells) of the volume launch plan (pad columns included)"""
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(self.L.b200fdtd_plan_info(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def energy(self):
        self._pre()
        e = C.c_double()
        check(self.L.b200fdtd_energy(self.h, C.byref(e)))
        return e.value

    def sync(self):
        check(self.L.b200fdtd_sync(self.h))

    @property
    def ts(self):
        v = C.c_int64()
        check(self.L.b200fdtd_get_timestep(self.h, C.byref(v)))
        return v.value

    def set_timestep(self, ts):
        check(self.L.b200fdtd_set_timestep(self.h, int(ts)))

    @property
    def num_samples(self):
        v = C.c_int()
        check(self.L.b200fdtd_num_samples(self.h, C.byref(v)))
        return v.value


def launch_count():
    return int(_lib.lib().b200fdtd_launch_count())


class FarfieldSources:
    """equivalent currents of one (frequency, centre) resident on the device: the reference asks for the far field once per
    phi (73 calls, …microstrip_3d.py:221-238); the currents are uploaded once and every call is one launch"""

    def __init__(self, pos, J, M, device=0):
        self.dev = torch.device("cuda", int(device))
        self.pos = torch.as_tensor(np.ascontiguousarray(pos, np.float32), device=self.dev)
        self.J = torch.as_tensor(np.ascontiguousarray(np.stack([np.real(J), np.imag(J)], -1), np.float32), device=self.dev)
        self.M = torch.as_tensor(np.ascontiguousarray(np.stack([np.real(M), np.imag(M)], -1), np.float32), device=self.dev)
        self.world, self.group = 1, None

    @classmethod
    def from_device(cls, pos, J, M):
        """pos [3][n], J/M [3][n][2] float32 CUDA tensors formed on the device (postproc.device_sources)"""
        o = cls.__new__(cls)
        o.dev, o.pos, o.J, o.M = pos.device, pos, J, M
        o.world, o.group = 1, None
        return o


def farfield(pos, J, M, k, theta, phi, device=0, sources=None):
    """K11 on device.  pos [3][n], J/M [3][n] complex, theta/phi radians ->
    (N_theta, N_phi, L_theta, L_phi) complex128 numpy arrays [ndir]."""
    L = _lib.lib()
    src = sources if sources is not None else FarfieldSources(pos, J, M, device)
    dev, pos_t, Jt, Mt = src.dev, src.pos, src.J, src.M
    n = pos_t.shape[1]
    th, ph = _np(theta, np.float64), _np(phi, np.float64)
    out = torch.zeros((len(th), 4, 2), dtype=torch.float32, device=dev)
    s = torch.cuda.current_stream(dev)
    if n > 0:                                      # a z-slab rank may hold no part of the box
        check(L.b200fdtd_farfield(dev.index, C.c_void_p(s.cuda_stream), n, pos_t.data_ptr(), Jt.data_ptr(), Mt.data_ptr(),
                                  float(k), len(th), _ptr(th, c_d), _ptr(ph, c_d), out.data_ptr()))
    if getattr(src, "world", 1) > 1:               # the radiation integral is linear in the sources: add the ranks' sums
        o64 = out.to(torch.float64)
        torch.distributed.all_reduce(o64, group=src.group)
        out = o64
    o = out.cpu().numpy().astype(np.float64)
    oc = o[..., 0] + 1j * o[..., 1]
    return oc[:, 0], oc[:, 1], oc[:, 2], oc[:, 3]
