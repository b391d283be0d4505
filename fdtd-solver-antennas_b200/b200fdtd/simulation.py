"""Simulation: what `FDTD.Run` does (antenna_sim/solver_fdtd_openems_microstrip_3d.py:214).

Builds the operator for this rank's z-slab, loads it into the CUDA engine, runs the time loop
(end criterion on the energy estimate, App. A7; progress lines the GUI recognises,
gui_app.py:493-495) and collects probe series / DFTs and the NF2FF face spectra.

Multi-GPU: one process per GPU (torch.distributed); the grid is cut into z-slabs, each half
step is followed by a one-plane halo exchange with the z-neighbours.
"""
from __future__ import annotations

import math
import time

import numpy as np
import torch

from .constants import C0
from .operator import OperatorBuilder, Setup, gauss_signal, BC_MUR, BC_PML  # noqa: F401


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


class LazyResults(dict):
    """dict whose expensive entries are produced on first access (e.g. the host copy of the NF2FF face spectra: the far
    field on the CUDA path is formed from the device-resident accumulators and never needs it)"""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._thunks = {}

    def lazy(self, key, fn):
        self._thunks[key] = fn

    def __missing__(self, key):
        if key in self._thunks:
            v = self._thunks.pop(key)()
            self[key] = v
            return v
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._thunks

    def get(self, key, default=None):
        return self[key] if key in self else default


def _round_up(n, m):
    return (n + m - 1) // m * m


def slab_ranges(nz, world):
    """contiguous, balanced z-slabs of node planes: [(K0, K1)] * world"""
    base, rem = divmod(nz, world)
    out, k = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((k, k + n))
        k += n
    return out


def _default_engine_factory(nx, ny, nz, px, device):
    from .engine import Engine
    return Engine(nx, ny, nz, px=px, device=device)


class Simulation:
    def __init__(self, setup: Setup, device=0, rank=0, world=1, group=None, engine_factory=None,
                 build_device=None, px_align=32, log=None, nf2ff_freqs=None, probe_freqs=None, align_x_slabs=True, compress_pml=True,
                 fused_multi=True, nf2ff_td=None):
        self.setup = setup
        self.rank, self.world, self.group = int(rank), int(world), group
        self.device = device
        self.log = log or (lambda *a: None)
        self.engine_factory = engine_factory or _default_engine_factory
        self.build_device = build_device       # None: chosen in prepare() from the slab size
        self.px_align = px_align
        self.align_x_slabs = bool(align_x_slabs)
        self.compress_pml = bool(compress_pml)
        self.fused_multi = bool(fused_multi)
        # time-domain store of the NF2FF face samples (far field at any frequency after the run): None = if it fits the
        # memory budget (B200FDTD_NF2FF_TD_GB, default 8 GB per GPU and at most 30 % of the free HBM), False = never
        self.nf2ff_td = nf2ff_td
        self.nf2ff_freqs = None if nf2ff_freqs is None else np.atleast_1d(np.asarray(nf2ff_freqs, np.float64))
        self.probe_freqs = None if probe_freqs is None else np.atleast_1d(np.asarray(probe_freqs, np.float64))
        self.engine = None
        self.results = None

    # ------------------------------------------------------------------ set-up
    def prepare(self):
        s = self.setup
        t0 = time.time()
        nz_glob = len(s.lines[2])
        self.slabs = slab_ranges(nz_glob, self.world)
        if min(b - a for a, b in self.slabs) < 2:
            raise ValueError(f"{nz_glob} z-planes are too few for {self.world} slabs")
        self.K0, self.K1 = self.slabs[self.rank]
        if self.build_device is None:
            # the operator build is torch tensor code: on the GPU for big slabs (0.9 s for 106 M cells), on the host for
            # small ones, where kernel-launch and sync overheads of thousands of tiny ops dominate (3 s vs 0.15 s at 1.3 M cells)
            slab_cells = len(s.lines[0]) * len(s.lines[1]) * (self.K1 - self.K0)
            use_gpu = torch.cuda.is_available() and slab_cells >= 4_000_000
            self.build_device = torch.device("cuda", self.device) if use_gpu else torch.device("cpu")
        B = self.builder = OperatorBuilder(s, device=self.build_device, k_nodes=(self.K0, self.K1))
        nx, ny, nz = B.n
        self.nx, self.ny, self.nz_glob = nx, ny, nz
        self.nz = self.K1 - self.K0
        self.px = _round_up(nx, self.px_align)
        dt = B.estimate_timestep() * s.timestep_factor
        if self.world > 1:
            t = torch.tensor([dt], dtype=torch.float64)
            if torch.distributed.get_backend(self.group) == "nccl":
                t = t.cuda(self.device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN, group=self.group)
            dt = float(t.item())
        self.dt = dt
        self.signal, self.exc_len = gauss_signal(s.f0, s.fc, dt, s.nrts) if s.f0 > 0 else (np.zeros(1, np.float32), 0)
        self.nyquist = B.nyquist(dt)
        self.interval = max(1, self.nyquist // max(1, s.oversampling))
        E = self.engine = self.engine_factory(nx, ny, self.nz, self.px, self.device)
        vv, vi, ii, iv = B.coefficients(self.K0, self.K1, self.px, dt)
        E.set_coeffs(vv, vi, ii, iv)
        del vv, vi, ii, iv
        self.compression = {}
        if B.row_compression is not None and hasattr(E, "set_row_compression"):
            for which in (0, 1):
                xv, meta = B.row_compression[which]
                if xv is not None:
                    self.compression[which] = E.set_row_compression(which, xv, meta) + (int(xv.shape[0]),)
        lin = lambda c, k, j, i: ((c * (self.nz + 2) + (k - self.K0 + 1)) * ny + j) * self.px + i   # noqa: E731
        own = lambda k: (k >= self.K0) & (k < self.K1)                                                  # noqa: E731
        # excitation
        comp, i, j, k, amp, delay = B.excitation_list(dt)
        sel = own(k)
        self.n_exc = int(sel.sum())
        if self.n_exc:
            E.set_excitation(lin(comp[sel], k[sel], j[sel], i[sel]), amp[sel], delay[sel], self.signal)
        # Mur
        mur = B.mur_list(dt)
        self.n_mur = 0
        if mur is not None:
            comp, dst, src, coeff = mur
            sel = own(dst[2])
            self.n_mur = int(sel.sum())
            if self.n_mur:
                # the inward neighbour of a z-face edge lives in the same slab (checked by the builder)
                d_lin = lin(comp[sel], dst[2][sel], dst[1][sel], dst[0][sel])
                s_lin = lin(comp[sel], src[2][sel], src[1][sel], src[0][sel])
                # face by face in memory order: rows of a z- or y-face become runs of stride 1, columns of an x-face runs
                # of stride px (the engine stores such runs as segments instead of index lists); every edge appears
                # once, so the order of the list does not change any result
                o = np.lexsort((d_lin, B.mur_face[sel]))
                E.set_mur(d_lin[o], s_lin[o], coeff[sel][o])
        # PML
        boxes = []
        self.pml_cells = 0
        for (ri, rj, rk) in B.pml_boxes():
            k0, k1 = max(rk[0], self.K0), min(rk[1], self.K1)
            if k1 <= k0:
                continue
            full_rows = ri[0] == 0 and ri[1] == nx
            if not full_rows and self.align_x_slabs:
                # narrow x-slab widened to float4-aligned columns (the extra columns get the identity coefficients
                # a = fo = fn = 1 that the formulas give outside the PML) so a narrow-slab launch of the volume kernel
                # (csrc/b200fdtd.cu MODE 2) can own these columns instead of separate pre/post passes
                ri = (0, min(nx, _round_up(ri[1], 4))) if ri[0] == 0 else ((ri[0] // 4) * 4, nx)
            bx = ri[1] - ri[0]
            if full_rows:
                bxp = self.px              # whole x-rows: padded to the row pitch so the volume kernels can fuse this slab
            elif self.align_x_slabs:
                bxp = _round_up(bx, 4)     # pad columns of the grid (i >= nx) or identity columns
            else:
                bxp = bx
            fusable = full_rows or self.align_x_slabs
            co = B.pml_coefficients((ri, rj, (k0, k1)), dt, pad_to=bxp, compress=fusable and self.compress_pml)
            box = dict(x0=ri[0], y0=rj[0], z0=k0 - self.K0, bx=bxp, by=rj[1] - rj[0], bz=k1 - k0)
            box.update(co)
            self.pml_cells += box["bx"] * box["by"] * box["bz"]
            boxes.append(box)
        if boxes:
            E.set_pml(boxes)
        # probes
        self.nrts_sized = s.nrts
        self.max_samples = s.nrts // self.interval + 2
        self.probe_names, kinds, offs, idxs, ws = [], [], [0], [], []
        for (name, kind, comp, i, j, k, w) in B.probe_lists():
            sel = own(k)
            self.probe_names.append(name); kinds.append(kind)
            idxs.append(lin(comp[sel], k[sel], j[sel], i[sel])); ws.append(w[sel])
            offs.append(offs[-1] + int(sel.sum()))
        if self.probe_freqs is None:
            self.probe_freqs = s.probe_freqs if s.probe_freqs is not None else np.zeros(0)
        if kinds:
            E.set_probes(kinds, offs, np.concatenate(idxs) if idxs else np.zeros(0, np.int64),
                         np.concatenate(ws) if ws else np.zeros(0), self.interval, self.max_samples, self.probe_freqs, dt)
        # NF2FF
        self.faces = B.nf2ff_faces()
        self.local_faces = []
        if self.faces:
            if self.nf2ff_freqs is None:
                fr = s.nf2ff.get("frequency") if s.nf2ff else None
                self.nf2ff_freqs = np.atleast_1d(np.asarray(fr if fr is not None else [s.f0], np.float64))
            loc = []
            for fi, F in enumerate(self.faces):
                n = F["normal"]; a, b = (n + 1) % 3, (n + 2) % 3
                L = dict(normal=n, plane=F["plane"], a0=F["a0"], a1=F["a1"], b0=F["b0"], b1=F["b1"])
                if n == 2:
                    if not own(np.array(F["plane"])):
                        continue
                    L["plane"] = F["plane"] - self.K0
                else:
                    key0, key1 = ("a0", "a1") if a == 2 else ("b0", "b1")
                    z0, z1 = max(F[key0], self.K0), min(F[key1], self.K1 - 1)
                    if z1 < z0:
                        continue
                    L[key0], L[key1] = z0 - self.K0, z1 - self.K0
                    L["z_off"] = z0 - F[key0]
                L["face"] = fi
                loc.append(L)
            self.local_faces = loc
            if loc:
                il = [1.0 / B.len_p[0], 1.0 / B.len_p[1], self._z_slice(1.0 / B.len_p[2])]
                idl = [1.0 / B.len_d[0], 1.0 / B.len_d[1], self._z_slice(1.0 / B.len_d[2])]
                E.set_nf2ff(loc, self.nf2ff_freqs, self.interval, dt, il, idl)
            self._setup_td_store()
        self.prepare_s = time.time() - t0
        self.cells = nx * ny * nz
        return self

    def _setup_td_store(self):
        """openEMS keeps the face samples themselves (HDF5 dumps) and CalcNF2FF transforms them at whatever frequency the caller
        names (…microstrip_3d.py:225).  Here the samples stay in HBM when they fit; all ranks decide alike."""
        import os
        E = self.engine
        self.td_bytes = 0
        want = self.nf2ff_td is not False and hasattr(E, "set_nf2ff_td")
        nodes = sum((L["a1"] - L["a0"] + 1) * (L["b1"] - L["b0"] + 1) for L in self.local_faces)
        need = nodes * 16 * self.max_samples
        ok = want
        if want and self.nf2ff_td is None:
            budget = float(os.environ.get("B200FDTD_NF2FF_TD_GB", "8")) * 2 ** 30
            if torch.cuda.is_available() and isinstance(getattr(E, "device", None), torch.device):
                budget = min(budget, 0.3 * torch.cuda.mem_get_info(E.device)[0])
            ok = need <= budget
        if self.world > 1:
            t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64)
            if torch.distributed.get_backend(self.group) == "nccl":
                t = t.cuda(self.device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN, group=self.group)
            ok = bool(t.item() > 0.5)
        if ok and self.local_faces:
            self.td_bytes = E.set_nf2ff_td(self.max_samples)
        self.td_store = bool(ok)

    def _assemble_faces(self, local, nf):
        """local per-face arrays [4][nf][nb_loc][na_loc][2] (this slab's part) -> global complex faces, summed over the slabs"""
        full = []
        for F in self.faces:
            na, nb = F["a1"] - F["a0"] + 1, F["b1"] - F["b0"] + 1
            full.append(np.zeros((4, nf, nb, na), np.complex128))
        for q, L in enumerate(self.local_faces):
            acc = _np(local[q]).astype(np.float64)
            acc = acc[..., 0] + 1j * acc[..., 1]
            F = self.faces[L["face"]]
            n = F["normal"]; a = (n + 1) % 3
            if n == 2:
                full[L["face"]][...] = acc
            elif a == 2:      # a-axis is z
                o = L["z_off"]; full[L["face"]][:, :, :, o:o + acc.shape[3]] = acc
            else:             # b-axis is z
                o = L["z_off"]; full[L["face"]][:, :, o:o + acc.shape[2], :] = acc
        if self.world > 1:
            full = [self._allreduce_np(np.stack([f.real, f.imag], -1)) for f in full]
            full = [f[..., 0] + 1j * f[..., 1] for f in full]
        scale = 2.0 * self.interval * self.dt          # single-sided pulse spectrum, like DFT_time2freq (App. A5/A6)
        return [f * scale for f in full]

    def _device_faces_ok(self):
        E = self.engine
        return (isinstance(getattr(E, "device", None), torch.device) and E.device.type == "cuda"
                and not getattr(self.builder, "nf2ff_mirrors", [])
                and (self.world == 1 or torch.distributed.get_backend(self.group) == "nccl"))

    def nf2ff_spectra(self, freqs):
        """face spectra at arbitrary frequencies from the stored time-domain samples (collective on z-slab runs)"""
        if not getattr(self, "td_store", False):
            raise ValueError("the NF2FF face samples were not kept (they did not fit the memory budget B200FDTD_NF2FF_TD_GB); "
                             "register the frequency before the run: CreateNF2FFBox(frequency=[...]) or FDTD.nf2ff_freqs")
        freqs = np.atleast_1d(np.asarray(freqs, np.float64))
        ns = self.results["n_samples"]
        local = self.engine.nf2ff_td_dft(freqs, ns) if self.local_faces else []
        return self._assemble_faces(local, len(freqs))

    def _z_slice(self, arr):
        """global per-line array -> local array with nz+2 entries (entry 0 = plane K0-1)"""
        out = np.ones(self.nz + 2)
        lo, hi = self.K0 - 1, self.K1 + 1
        a, b = max(lo, 0), min(hi, len(arr))
        out[a - lo:b - lo] = arr[a:b]
        return out

    # ------------------------------------------------------------------ compressed operator on the host
    def export_operator(self, pin=True):
        """Host copy of the operator in its compressed form: per pass the x-vector table, the 32-byte row records and the
        rows that are not of the form scale*xvec (RowFactor).  ~0.1 % of the 48 B/cell of the full arrays."""
        E = self.engine
        out = {}
        for which, (ca, cb) in enumerate(((E.vv, E.vi), (E.ii, E.iv))):
            xv, meta = E._keep[f"cmp{which}"]                # device tensors after the device-side verification
            ids = meta.view(self.nz + 2, self.ny, 32)[..., 24:30]
            rows, slots = [], []
            for slot in range(6):
                arr = (ca, cb)[slot // 3][slot % 3]
                full = (ids[..., slot] == 255) & (arr.abs().amax(-1) > 0)
                idx = full.nonzero()
                slots.append(idx)
                rows.append(arr[idx[:, 0], idx[:, 1]])
            H = dict(xvecs=xv.cpu(), meta=meta.cpu(), full_idx=[i.cpu() for i in slots], full_rows=[r.cpu() for r in rows])
            if pin:
                H = {k: ([t.pin_memory() for t in v] if isinstance(v, list) else v.pin_memory()) for k, v in H.items()}
            out[which] = H
        return out

    @staticmethod
    def operator_nbytes(op):
        n = 0
        for H in op.values():
            for v in H.values():
                n += sum(t.numel() * t.element_size() for t in v) if isinstance(v, list) else v.numel() * v.element_size()
        return n

    def load_operator(self, op):
        """H2D of a compressed operator, expansion into the bound full arrays on the device, and (re)verification.
        Returns (row-slots compressed, row-slots demoted) per pass."""
        E = self.engine
        dev = E.device
        res = {}
        for which, (ca, cb) in enumerate(((E.vv, E.vi), (E.ii, E.iv))):
            H = op[which]
            xv = H["xvecs"].to(dev, non_blocking=True)
            meta = H["meta"].to(dev, non_blocking=True)
            xv = xv.contiguous(); meta = meta.contiguous()
            E.expand_rows(which, xv, meta)                     # C-ABI b200fdtd_expand_rows: one kernel, HBM write speed
            for slot in range(6):
                arr = (ca, cb)[slot // 3][slot % 3]
                idx = H["full_idx"][slot].to(dev, non_blocking=True)
                if idx.numel():
                    arr[idx[:, 0], idx[:, 1]] = H["full_rows"][slot].to(dev, non_blocking=True)
            res[which] = E.set_row_compression(which, xv, meta)
        return res

    def keep_host_operator(self):
        """FDTD.Run(setup_only=True): keep the operator on the HOST in the form the host builder produces it (pinned), so a later
        Run of the same scene ships it to the device instead of rebuilding it"""
        if hasattr(self.engine, "expand_rows") and all(f"cmp{w}" in self.engine._keep for w in (0, 1)):
            self.host_op = self.export_operator(pin=True)
        return self

    def restart(self):
        """a new run of the prepared scene: operator host -> device (if a host copy is kept), all state back to zero"""
        E = self.engine
        t0 = time.time()
        self.reloaded = None
        if getattr(self, "host_op", None) is not None:
            self.reloaded = self.load_operator(self.host_op)
        E.reset_state()
        self.restart_s = time.time() - t0
        if getattr(self, "_tv", None) is not None:
            self._vcur = self._ccur = 0
        self.results = None
        return self

    # ------------------------------------------------------------------ halo exchange (z-slabs)
    # Data dependence: the E update of my plane 0 reads (Hx,Hy) of the lower neighbour's top plane;
    # the H update of my top plane reads (Ex,Ey) of the upper neighbour's plane 0.  Planes are sent straight out of /
    # received straight into the field arrays (each component plane is contiguous: no staging copies).  The NCCL
    # transfers run on the process group's stream while the interior launch of the next half step runs on the
    # engine stream; only the one boundary plane waits for them (b200fdtd_half_step_part).
    def _views(self):
        if getattr(self, "_tv", None) is None:
            E = self.engine
            self._tv = torch.from_numpy(E.volt) if isinstance(E.volt, np.ndarray) else E.volt
            self._tc = torch.from_numpy(E.curr) if isinstance(E.curr, np.ndarray) else E.curr
            self._gpu = self._tv.is_cuda
            self._pend_e, self._pend_h = [], []
            # fused H->E steps (second field copy, C-ABI b200fdtd_fused_step_part): CUDA engine with >= 3 planes per slab
            # (the CPU oracle engine of the tests emulates the protocol so that gloo runs cover this host logic too)
            self._fused = bool((self._gpu or getattr(E, "emulates_fused_steps", False)) and self.fused_multi
                               and hasattr(E, "fused_step_part") and self.nz >= 3)
            self._copies_v, self._copies_c = [self._tv], [self._tc]
            if self._fused:
                # b200fdtd_fused_step_part needs every PML box inside the volume launches (no separate pre/post boxes)
                if hasattr(E, "plan_info") and E.plan_info()[2] != 0:
                    self._fused = False
            if self._fused:
                try:
                    E.bind_alt_fields()
                    as_t = lambda a: torch.from_numpy(a) if isinstance(a, np.ndarray) else a      # noqa: E731
                    self._copies_v.append(as_t(E.volt2)); self._copies_c.append(as_t(E.curr2))
                except Exception:            # e.g. out of memory: keep the separate half steps
                    self._fused = False
            self._vcur = self._ccur = 0
            # the fused launch in two halves hides the wait for the upper ghost E behind the lower half; B200FDTD_SPLIT_HE=0:
            # one launch (the exchange then has to fit behind the boundary-plane launches of part 0)
            import os
            self._split_he = os.environ.get("B200FDTD_SPLIT_HE", "1") != "0"
        return self._tv, self._tc

    def _cur(self):
        """the copies that hold E and H right now (the fused steps ping-pong between two copies)"""
        return self._copies_v[self._vcur], self._copies_c[self._ccur]

    def _exchange_async(self, f, direction, ncomp):
        """direction +1: my top owned plane -> upper neighbour's lower ghost; -1: my plane 0 -> lower neighbour's upper ghost.
        The P2P op lists (plane views of a given array) are built once and reused: at ~1 ms of GPU work per step the host
        side of the exchange is on the critical path."""
        dist = torch.distributed
        cache = self.__dict__.setdefault("_p2p_cache", {})
        key = (f.data_ptr(), direction, ncomp)
        ops = cache.get(key)
        if ops is None:
            ops = []
            up, dn = self.rank + 1, self.rank - 1
            if direction > 0:
                if up < self.world:
                    ops += [dist.P2POp(dist.isend, f[c, self.nz], up, self.group) for c in range(ncomp)]
                if dn >= 0:
                    ops += [dist.P2POp(dist.irecv, f[c, 0], dn, self.group) for c in range(ncomp)]
            else:
                if dn >= 0:
                    ops += [dist.P2POp(dist.isend, f[c, 1], dn, self.group) for c in range(ncomp)]
                if up < self.world:
                    ops += [dist.P2POp(dist.irecv, f[c, self.nz + 1], up, self.group) for c in range(ncomp)]
            cache[key] = ops
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def _wait(works):
        for w in works:
            w.wait()            # NCCL: the current (engine) stream waits on the device; gloo: host wait

    def _e_half(self):
        """E half step in two parts around the wait for the lower ghost H; then my plane 0 goes down"""
        E = self.engine
        E.half_step_part(0, 0)                       # E: pre passes + planes [1,nz)      (overlaps the H halo)
        self._wait(self._pend_h); self._pend_h = []
        E.half_step_part(0, 1)                       # E: plane 0 + post passes
        self._pend_e = self._exchange_async(self._cur()[0], -1, 2)

    def _h_half(self):
        E = self.engine
        E.half_step_part(1, 0)                       # H: pre passes + planes [0,nz-1)    (overlaps the E halo)
        self._wait(self._pend_e); self._pend_e = []
        E.half_step_part(1, 1)                       # H: top plane + post passes, ++ts
        self._pend_h = self._exchange_async(self._cur()[1], +1, 2)

    def _fused_step(self):
        """H(n) + E(n+1): boundary planes by separate launches, interior planes by the fused launch (csrc, part 0..3)"""
        E = self.engine
        E.fused_step_part(0)                         # H: plane 0 + interior PML slabs
        if self._split_he:
            E.fused_step_part(4)                     # fused launch, lower half of the interior planes (overlaps the E halo)
        self._wait(self._pend_e); self._pend_e = []
        E.fused_step_part(1)                         # H: top plane (needed the upper ghost E)
        h_new = self._copies_c[self._ccur ^ 1]       # the H launches wrote the other copy; it becomes current in part 2
        self._pend_h = self._exchange_async(h_new, +1, 2)
        E.fused_step_part(2)                         # fused launch, upper half (overlaps the H halo), ++ts
        self._ccur ^= 1
        self._wait(self._pend_h); self._pend_h = []
        E.fused_step_part(3)                         # E: interior PML slabs, plane 0 (lower ghost H_new), top plane; excitation
        self._vcur ^= 1
        self._pend_e = self._exchange_async(self._cur()[0], -1, 2)

    def _sample_multi(self, pipelined):
        """probe / NF2FF sampling of a z-slab rank at a sampling step.  pipelined: between two fused steps the E update of the
        next step is already done; the sample takes E from the copy that fused launch only read (csrc: launch_sampling)."""
        E = self.engine
        self._wait(self._pend_h); self._pend_h = []
        if self.faces:                               # NF2FF node interpolation reads E of plane K0-1 too
            ev = self._copies_v[self._vcur ^ 1] if pipelined else self._cur()[0]
            if pipelined:                            # (the E exchange of the running step must not be reordered around it)
                self._wait(self._pend_e); self._pend_e = []
            self._wait(self._exchange_async(ev, +1, 3))
        E.half_step_raw(3 if pipelined else 2)

    def _step_multi(self, n):
        """n steps of a z-slab rank.  With the fused steps the whole call is one span E(0) | H(0)+E(1) | ... | H(n-1) (an even
        number of fused steps, so the state ends in the bound arrays), sampling points inside it are served in the pipelined
        state; without them: E and H half steps with one exchange each."""
        import contextlib
        E = self.engine
        self._views()
        if self._gpu:
            # everything torch did on its current stream so far (zero fills of the second field copy, of the probe series and
            # of the NF2FF accumulators, operator uploads) is ordered before the first launch on the engine stream
            E._pre()
        ctx = torch.cuda.stream(E.stream) if self._gpu else contextlib.nullcontext()
        sampling = bool(self.faces) or bool(self.probe_names)
        iv = self.interval
        with ctx:
            fused = (n - 2 if (n - 1) & 1 else n - 1) if self._fused else 0
            done = 0
            while done < n:
                self._e_half()
                if done == 0:
                    for _ in range(fused):
                        self._fused_step()
                        if sampling and (E.ts % iv) == 0:
                            self._sample_multi(True)
                    done += fused
                self._h_half()
                done += 1
                if sampling and (E.ts % iv) == 0:
                    self._sample_multi(False)
            self._wait(self._pend_h); self._pend_h = []
            self._wait(self._pend_e); self._pend_e = []
            if self._fused and (self._vcur or self._ccur):
                raise RuntimeError("fused z-slab steps did not return to the bound field arrays")
            if self._fused:
                E.reset_current_copy()               # (the library's PML flux copy goes back to the caller's arrays too)
        if self._gpu:
            torch.cuda.current_stream(E.device).wait_stream(E.stream)

    # ------------------------------------------------------------------ time loop
    def energy(self):
        e = self.engine.energy()
        if self.world > 1:
            t = torch.tensor([e], dtype=torch.float64)
            if torch.distributed.get_backend(self.group) == "nccl":
                t = t.cuda(self.device)
            torch.distributed.all_reduce(t, group=self.group)
            e = float(t.item())
        return e

    def run(self, nrts=None, end_criteria=None, verbose=0, check_every=None, use_graph=True):
        s = self.setup
        nrts = int(s.nrts if nrts is None else nrts)
        end_criteria = float(s.end_criteria if end_criteria is None else end_criteria)
        E = self.engine
        if check_every is None:
            # a multiple of the sampling interval close to ~1% of the run
            check_every = self.interval * max(1, int(round(max(50, nrts / 100) / self.interval)))
        e_max, e_now = 0.0, 0.0
        t0 = time.time(); t_last = t0; ts_last = 0
        stop_reason = "NrTS"
        while E.ts < nrts:
            n = min(check_every, nrts - E.ts)
            if self.world > 1:
                self._step_multi(n)
            else:
                E.run(n, use_graph=use_graph)
            e_now = self.energy()
            if not math.isfinite(e_now):
                raise FloatingPointError(f"field energy is not finite at timestep {E.ts} (unstable time step?)")
            e_max = max(e_max, e_now)
            now = time.time()
            if verbose and self.rank == 0:
                speed = self.cells * (E.ts - ts_last) / max(now - t_last, 1e-9) / 1e6
                db = 10.0 * math.log10(e_now / e_max) if e_now > 0 and e_max > 0 else -200.0
                self.log(f"[@ {now - t0:8.2f}s] Timestep: {E.ts:>8d} || Speed: {speed:8.1f} MC/s "
                         f"({(now - t_last) / max(E.ts - ts_last, 1):.3e}s/TS) || Energy: ~{e_now:.2e} ({db:6.2f}dB)")
            t_last, ts_last = now, E.ts
            if e_max > 0 and e_now / e_max < end_criteria and E.ts > self.exc_len:
                stop_reason = "EndCriteria"
                break
        self.wall_s = time.time() - t0
        self.stop_reason = stop_reason
        self.timesteps = E.ts
        t1 = time.time()
        self.collect()
        self.collect_s = time.time() - t1
        return self

    # ------------------------------------------------------------------ results
    def collect(self):
        """bring the (small) results to the host; sums partial probe values over the slabs"""
        E = self.engine
        ns = E.num_samples if hasattr(E, "num_samples") else E.ts // self.interval
        if not isinstance(ns, int):
            ns = int(ns)
        ns = min(ns, E.ts // self.interval)
        res = {"dt": self.dt, "interval": self.interval, "timesteps": E.ts, "n_samples": ns, "probes": {}}
        if self.probe_names:
            series = _np(E.series)[:, :ns].astype(np.float64)
            dft = _np(E.probe_dft).astype(np.float64)
            if self.world > 1:
                series = self._allreduce_np(series); dft = self._allreduce_np(dft)
            k_list = [pl[1] for pl in self.builder.probe_lists()]
            for p, name in enumerate(self.probe_names):
                kind = k_list[p]
                t = (np.arange(1, ns + 1) * self.interval + (0.5 if kind == 1 else 0.0)) * self.dt
                res["probes"][name] = dict(kind=kind, t=t, val=series[p],
                                           dft=(dft[p, :, 0] + 1j * dft[p, :, 1]) if len(self.probe_freqs) else None)
            res["probe_freqs"] = self.probe_freqs
        if self.faces:
            nf = LazyResults(faces=self.faces, freqs=self.nf2ff_freqs, mirrors=list(getattr(self.builder, "nf2ff_mirrors", [])),
                             weights=[self.builder.face_weights(F) for F in self.faces],
                             spectra_fn=self.nf2ff_spectra if getattr(self, "td_store", False) else None, extra={}, sources={})
            dev_ok = self._device_faces_ok()
            if self.world > 1 and not dev_ok:
                nf["acc"] = self._assemble_faces(E.face_acc, len(self.nf2ff_freqs))      # collective: every rank takes part now
            else:
                # host copy of the spectra only if somebody asks for it (z-slab runs: a collective, every rank must ask)
                nf.lazy("acc", lambda: self._assemble_faces(E.face_acc, len(self.nf2ff_freqs)))
            if dev_ok:
                # CUDA engine, no image walls: the equivalent currents are formed on the device from this rank's part of
                # the faces (postproc.device_sources); z-slab ranks add their far-field sums (a few numbers per direction)
                loc = []
                for q, L in enumerate(self.local_faces):
                    F = self.faces[L["face"]]
                    xa, xb, wa, wb = nf["weights"][L["face"]]
                    n = F["normal"]; a = (n + 1) % 3
                    if n != 2:
                        o = L["z_off"]
                        if a == 2:
                            m = L["a1"] - L["a0"] + 1; xa, wa = xa[o:o + m], wa[o:o + m]
                        else:
                            m = L["b1"] - L["b0"] + 1; xb, wb = xb[o:o + m], wb[o:o + m]
                    loc.append(dict(normal=n, side=F["side"], coord=F["coord"], xa=xa, xb=xb, wa=wa, wb=wb, acc=E.face_acc[q]))
                nf["device"] = dict(faces=loc, scale=2.0 * self.interval * self.dt, dev=E.device, world=self.world, group=self.group,
                                    td_dft=(lambda f: self.engine.nf2ff_td_dft([f], self.results["n_samples"])) if getattr(self, "td_store", False) else None)
            res["nf2ff"] = nf
        self.results = res
        return res

    def _allreduce_np(self, a):
        t = torch.from_numpy(np.ascontiguousarray(a, np.float64))
        if torch.distributed.get_backend(self.group) == "nccl":
            t = t.cuda(self.device)
        torch.distributed.all_reduce(t, group=self.group)
        return t.cpu().numpy()
