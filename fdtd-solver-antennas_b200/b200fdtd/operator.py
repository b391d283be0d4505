"""Host-side operator build: CSX scene -> engine-level inputs (include/b200fdtd.h).

What openEMS does between `FDTD.Run` being called and the first time step
(antenna_sim/solver_fdtd_openems_microstrip_3d.py:214): voxelise the primitives on the
rectilinear grid, turn materials into edge capacitances/conductances and face inductances,
derive the vv/vi/ii/iv update coefficients, mark PEC edges, place lumped
resistors, excitations, probes, Mur/PML boundary data and the NF2FF box.

Volume arrays are produced per z-range so a z-slab rank only ever materialises its own slab.
The arithmetic is float64 torch (CPU or CUDA tensor ops), stored as float32 like openEMS
(`FDTD_FLOAT=float`).  This is set-up work, not the hot path; the hot path is csrc/b200fdtd.cu.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np
import torch

from .constants import C0, EPS0, MUE0

BC_PEC, BC_PMC, BC_MUR, BC_PML = 0, 1, 2, 3
PML_GRADE = 2.5           # geometric grading base of the default openEMS UPML profile (App. A4)
PML_R0 = 1e-6             # theoretical reflection of the profile


def parse_bc(bc):
    """['MUR']*6 | ['PML_8']*6 | [3]*6 ... -> (types[6], pml_cells[6])  """
    if len(bc) != 6:
        raise ValueError("SetBoundaryCond needs 6 entries (xmin,xmax,ymin,ymax,zmin,zmax)")
    types, cells = [], []
    for b in bc:
        if isinstance(b, str):
            s = b.upper()
            if s == "PEC":
                t, c = BC_PEC, 0
            elif s == "PMC":
                t, c = BC_PMC, 0
            elif s == "MUR":
                t, c = BC_MUR, 0
            elif s.startswith("PML_"):
                t, c = BC_PML, int(s[4:])
            else:
                raise ValueError(f"unknown boundary condition '{b}'")
        else:
            t = int(b)
            if t not in (0, 1, 2, 3):
                raise ValueError(f"unknown boundary condition {b}")
            c = 8 if t == BC_PML else 0
        if t == BC_PMC:
            raise NotImplementedError("PMC boundaries are outside the reference's path")
        if t == BC_PML and not (4 <= c <= 50):
            raise ValueError("PML size must be in 4..50 cells")
        types.append(t); cells.append(c)
    return types, cells


def gauss_signal(f0, fc, dt, nrts):
    """openEMS Gaussian pulse (App. A2): s(t)=cos(2 pi f0 (t-t0)) exp(-(2 pi fc t/3 - 3)^2), t0 = 9/(2 pi fc).
    Returns (signal f32 with signal[0] = 0 and signal[n] = s((n-1) dt), length in steps)."""
    t0 = 9.0 / (2.0 * math.pi * fc)
    length = int(2.0 * t0 / dt)
    if length > nrts:
        length = int(nrts)
    t = np.arange(length, dtype=np.float64) * dt
    s = np.cos(2.0 * math.pi * f0 * (t - t0)) * np.exp(-(2.0 * math.pi * fc * t / 3.0 - 3.0) ** 2)
    return np.concatenate([[0.0], s]).astype(np.float32), length


@dataclass
class Setup:
    """Everything `Run` needs, in SI units, independent of the CSXCAD object model."""
    lines: list                      # 3 arrays of mesh lines [m], sorted unique
    bc: list                         # 6 boundary types
    pml_cells: list                  # 6 PML sizes
    materials: list = field(default_factory=list)   # dicts: eps, kappa, priority, order, prim (contains(pts), bbox)
    metals: list = field(default_factory=list)      # dicts: priority, order, prim
    lumped: list = field(default_factory=list)      # dicts: ny, R, caps, lo[3], hi[3] (m)
    excitations: list = field(default_factory=list)  # dicts: vec[3], delay(s), lo[3], hi[3], prim
    probes: list = field(default_factory=list)      # dicts: name, p_type, weight, norm_dir, start[3], stop[3]
    nf2ff: dict | None = None        # dict: start[3], stop[3] (m), frequency list
    f0: float = 0.0
    fc: float = 0.0
    nrts: int = 0
    end_criteria: float = 1e-5
    timestep_factor: float = 1.0
    oversampling: int = 4
    probe_freqs: np.ndarray | None = None


def _snap(lines, v):
    return int(np.argmin(np.abs(lines - v)))


class RowFactor:
    """Row compression of one pass (E: vv,vi / H: ii,iv), host side of b200fdtd_set_row_compression.

    On a rectilinear mesh most coefficient rows are scale(j,k) * xvec[i] for a handful of x-vectors.  The float32 operator
    is therefore DEFINED row by row in factored form: with i* the first non-zero column of the float64 row,
        coefficient[i] = fl32( fl32(row[i*]) * fl32(row[i] / row[i*]) )
    (at most 1.5 ulp from rounding the float64 value directly, exact at i*).  The definition depends on the row alone -
    not on which rows share an x-vector, on the chunking of the build or on how the grid is cut into z-slabs - so a
    z-slab rank builds bit for bit the rows the single-GPU run builds.  Compression is then a lossless encoding on top:
    rows whose normalised float32 vector fl32(row / row[i*]) equals a table entry bit for bit are recorded as
    (scale, id); the CUDA kernels rebuild them with one multiply and stay bit-identical to a run from the full arrays."""
    MAX_VECS = 250

    def __init__(self, nx, px, nzp, ny, dev, nslots=6, rec_bytes=32):
        """nslots scales + nslots ids per row record of rec_bytes (operator passes: 6 / 32; PML slabs: 9 / 48)"""
        self.nx, self.px, self.dev = nx, px, dev
        self.nslots, self.id0 = nslots, 4 * nslots
        self.vec32 = []                                   # shared table of this pass: float32 [px] each
        self.slots = [[] for _ in range(nslots)]          # per slot: ids of the table entries rows of this slot have used
        self.meta_f = torch.zeros((nzp, ny, rec_bytes // 4), dtype=torch.float32, device=dev)
        self.meta_u8 = self.meta_f.view(torch.uint8)      # [nzp, ny, rec_bytes]
        self.meta_u8[:, :, self.id0:self.id0 + nslots] = 255
        self.rows_total = 0
        self.rows_compressed = 0

    def _match(self, x32, gid, ids, unmatched):
        ok = unmatched & (x32 == self.vec32[gid][:self.nx]).all(-1)
        ids[ok] = gid
        unmatched &= ~ok

    def factor(self, arr, slot, plane0):
        """arr float64 [nk, ny, nx] -> float32 [nk, ny, nx]; records (scale, id) of the rows in meta[plane0:plane0+nk]"""
        nk, ny, nx = arr.shape
        nzm = arr != 0
        first = nzm.to(torch.int8).argmax(-1, keepdim=True)              # i*: first non-zero column (0 for an all-zero row)
        s = arr.gather(-1, first)                                         # row[i*]  (0 for an all-zero row)
        live = s != 0
        x32 = torch.where(live, arr / torch.where(live, s, torch.ones_like(s)), torch.zeros_like(arr)).to(torch.float32)
        s32 = s.to(torch.float32)
        out = s32 * x32                                                   # float32 multiply, IEEE rn: THE coefficient
        live = live.squeeze(-1)
        ids = torch.full((nk, ny), 255, dtype=torch.int64, device=self.dev)
        unmatched = live.clone()
        for gid in self.slots[slot]:
            self._match(x32, gid, ids, unmatched)
        for gid in range(len(self.vec32)):                                # vectors learnt by other slots of this pass
            if not bool(unmatched.any()):
                break
            if gid not in self.slots[slot]:
                n0 = int(unmatched.sum())
                self._match(x32, gid, ids, unmatched)
                if int(unmatched.sum()) != n0:
                    self.slots[slot].append(gid)
        for _ in range(3):                                # learn new x-vectors from rows nothing matched yet
            if not bool(unmatched.any()) or len(self.vec32) >= self.MAX_VECS:
                break
            idx = unmatched.nonzero()
            kk, jj = idx[len(idx) // 2].tolist()
            gid = len(self.vec32)
            v32 = torch.zeros(self.px, dtype=torch.float32, device=self.dev)
            v32[:nx] = x32[kk, jj]
            self.vec32.append(v32)
            self.slots[slot].append(gid)
            self._match(x32, gid, ids, unmatched)
        if len(self.vec32):
            # all-zero rows (edges that do not exist, PEC rows): scale 0 times a vector without negative entries is +0 everywhere
            pos = [g for g, v in enumerate(self.vec32) if bool((v >= 0).all())]
            if pos:
                ids[~live] = pos[0]
        self.meta_f[plane0:plane0 + nk, :, slot] = s32.squeeze(-1)
        self.meta_u8[plane0:plane0 + nk, :, self.id0 + slot] = ids.to(torch.uint8)
        self.rows_total += nk * ny
        self.rows_compressed += int((ids != 255).sum())
        return out

    def tables(self):
        if not self.vec32:
            return None, None
        return torch.stack(self.vec32).contiguous(), self.meta_u8.contiguous()


class OperatorBuilder:
    """Scene -> coefficient blocks and narrow-band index lists in GLOBAL grid indices."""

    def __init__(self, setup: Setup, device="cpu", k_nodes=None):
        self.s = setup
        self.dev = torch.device(device)
        self.x = [np.asarray(l, np.float64) for l in setup.lines]
        self.n = [len(l) for l in self.x]
        if min(self.n) < 3:
            raise ValueError("the mesh needs at least 3 lines per axis")
        self.d = [np.diff(l) for l in self.x]
        if any((d <= 0).any() for d in self.d):
            raise ValueError("mesh lines must be strictly increasing")
        # primal edge lengths (last entry repeats) and dual (node) widths (full cell at the ends)
        self.len_p = [np.concatenate([d, d[-1:]]) for d in self.d]
        self.len_d = [np.concatenate([d[:1], 0.5 * (d[:-1] + d[1:]), d[-1:]]) for d in self.d]
        self.h_lo = [np.concatenate([[0.0], 0.5 * d]) for d in self.d]
        self.h_hi = [np.concatenate([0.5 * d, [0.0]]) for d in self.d]
        self.mid = [0.5 * (l[:-1] + l[1:]) for l in self.x]
        nx, ny, nz = self.n
        # node planes owned by this builder (a z-slab; default: everything) and the cells they touch
        self.K0, self.K1 = (0, nz) if k_nodes is None else (int(k_nodes[0]), int(k_nodes[1]))
        if not (0 <= self.K0 < self.K1 <= nz):
            raise ValueError(f"bad z-slab {self.K0}:{self.K1} for {nz} planes")
        self.kc0, self.kc1 = max(0, self.K0 - 1), min(nz - 1, self.K1)
        self._pml_profiles()
        self._voxelise()
        self._mark_pec()
        self._lumped()
        self.dt = None

    # ------------------------------------------------------------------ geometry
    def _idx_range(self, coords, lo, hi, tol):
        """indices i with lo-tol <= coords[i] <= hi+tol  -> (i0, i1) half open"""
        i0 = int(np.searchsorted(coords, lo - tol, side="left"))
        i1 = int(np.searchsorted(coords, hi + tol, side="right"))
        return i0, i1

    def _tol(self):
        return 1e-9 * max(float(np.abs(l).max()) for l in self.x)

    def _voxelise(self):
        """cell-centre sampling of eps_r / kappa by priority (highest wins, later wins ties)"""
        nx, ny, nz = self.n
        nzc = self.kc1 - self.kc0
        eps = np.ones((nzc, ny - 1, nx - 1), np.float32)
        kap = np.zeros((nzc, ny - 1, nx - 1), np.float32)
        tol = self._tol()
        for m in sorted(self.s.materials, key=lambda m: (m["priority"], m["order"])):
            bb = m["prim"].bbox_m
            r = [self._idx_range(self.mid[a], bb[0][a], bb[1][a], tol) for a in range(3)]
            k0, k1 = max(r[2][0], self.kc0), min(r[2][1], self.kc1)
            if r[0][1] <= r[0][0] or r[1][1] <= r[1][0] or k1 <= k0:
                continue
            sl = (slice(k0 - self.kc0, k1 - self.kc0), slice(*r[1]), slice(*r[0]))
            if m["prim"].axis_aligned:
                eps[sl] = m["eps"]; kap[sl] = m["kappa"]
            else:
                Z, Y, X = np.meshgrid(self.mid[2][k0:k1], self.mid[1][slice(*r[1])], self.mid[0][slice(*r[0])], indexing="ij")
                inside = m["prim"].contains_m(np.stack([X, Y, Z], -1), tol)
                eps[sl][inside] = m["eps"]; kap[sl][inside] = m["kappa"]
        # zero-padded by one cell on every side: pad index p <-> cell p-1 (+kc0 along z)
        self.eps_pad = torch.zeros((nzc + 2, ny + 1, nx + 1), dtype=torch.float32, device=self.dev)
        self.kap_pad = torch.zeros_like(self.eps_pad)
        self.eps_pad[1:-1, 1:-1, 1:-1] = torch.from_numpy(eps).to(self.dev)
        self.kap_pad[1:-1, 1:-1, 1:-1] = torch.from_numpy(kap).to(self.dev)
        # cells outside this builder's z-range but inside the domain are not zero in reality; ranks ask only
        # for node planes whose cells they hold (kc range = node range widened by one), so this is never read.

    def _edge_boxes(self, prim, comp, tol):
        """index ranges (i0,i1),(j0,j1),(k0,k1) of comp-edges whose midpoint lies inside prim's bounding box"""
        bb = prim.bbox_m
        r = []
        for a in range(3):
            coords = self.mid[a] if a == comp else self.x[a]
            r.append(self._idx_range(coords, bb[0][a], bb[1][a], tol))
        return r

    def _mark_pec(self):
        """PEC edges: edge midpoint inside a metal primitive (closed box) — CalcPEC (App. A2)"""
        self.pec = []          # list of (comp, (i0,i1),(j0,j1),(k0,k1), mask or None)
        tol = self._tol()
        for m in self.s.metals:
            for comp in range(3):
                r = self._edge_boxes(m["prim"], comp, tol)
                if any(b <= a for a, b in r):
                    continue
                mask = None
                if not m["prim"].axis_aligned:
                    co = [self.mid[a][slice(*r[a])] if a == comp else self.x[a][slice(*r[a])] for a in range(3)]
                    Z, Y, X = np.meshgrid(co[2], co[1], co[0], indexing="ij")
                    mask = m["prim"].contains_m(np.stack([X, Y, Z], -1), tol)
                    if not mask.any():
                        continue
                self.pec.append((comp, r[0], r[1], r[2], mask))

    def _lumped(self):
        """lumped resistors (App. A5): per-edge conductance and PEC caps, in global indices"""
        self.lumped_G = []     # (comp, i, j, k, G) arrays
        for le in self.s.lumped:
            ny = le["ny"]
            nP, nPP = (ny + 1) % 3, (ny + 2) % 3
            lo_i = [_snap(self.x[a], le["lo"][a]) for a in range(3)]
            hi_i = [_snap(self.x[a], le["hi"][a]) for a in range(3)]
            if hi_i[ny] <= lo_i[ny]:
                raise ValueError("lumped element has no extent along its direction after snapping to the mesh")
            rng = [np.arange(lo_i[a], hi_i[a] + 1) for a in range(3)]
            rng[ny] = np.arange(lo_i[ny], hi_i[ny])              # edges along ny
            I = np.meshgrid(*rng, indexing="ij")
            idx = [g.ravel() for g in I]
            area = self.len_d[nP][idx[nP]] * self.len_d[nPP][idx[nPP]]
            length = self.len_p[ny][idx[ny]]
            # cross-section and length of the whole element (unit G = A_box / L_box)
            a_box = (self.len_d[nP][rng[nP]].sum()) * (self.len_d[nPP][rng[nPP]].sum())
            l_box = self.len_p[ny][rng[ny]].sum()
            R = float(le["R"])
            if R > 0:
                G = (area / length) / (a_box / l_box) / R
                self.lumped_G.append((ny, idx[0], idx[1], idx[2], G))
            else:
                self.pec.append((ny, (lo_i[0], hi_i[0] + (ny != 0)), (lo_i[1], hi_i[1] + (ny != 1)), (lo_i[2], hi_i[2] + (ny != 2)), None))
            if le.get("caps", True):
                for t in (nP, nPP):                              # transverse edges on both end planes
                    for plane in (lo_i[ny], hi_i[ny]):
                        r = [(lo_i[a], hi_i[a] + 1) for a in range(3)]
                        r[ny] = (plane, plane + 1)
                        r[t] = (lo_i[t], hi_i[t])                # edges along t between the box nodes
                        if r[t][1] > r[t][0]:
                            self.pec.append((t, r[0], r[1], r[2], None))

    # ------------------------------------------------------------------ PML profiles
    def _pml_profiles(self):
        """1-D grading g(D) [1/m] per axis at line and mid-cell positions; sigma_e/eps = sigma_m/mu = g*c (App. A4)"""
        self.g_line, self.g_mid, self.pml_lo, self.pml_hi = [], [], [], []
        for a in range(3):
            x, n = self.x[a], self.n[a]
            gl, gm = np.zeros(n), np.zeros(n)           # gm[i] describes the cell between line i and i+1
            lo_cells = self.s.pml_cells[2 * a] if self.s.bc[2 * a] == BC_PML else 0
            hi_cells = self.s.pml_cells[2 * a + 1] if self.s.bc[2 * a + 1] == BC_PML else 0
            if lo_cells + hi_cells + 1 >= n - 1:
                raise ValueError(f"PML ({lo_cells}+{hi_cells} cells) does not fit in {n} lines along axis {a}")
            for cells, side in ((lo_cells, 0), (hi_cells, 1)):
                if cells == 0:
                    continue
                if side == 0:
                    W = x[cells] - x[0]; Dl = x[cells] - x; Dm = x[cells] - self.mid[a]
                else:
                    W = x[-1] - x[n - 1 - cells]; Dl = x - x[n - 1 - cells]; Dm = self.mid[a] - x[n - 1 - cells]
                dl = W / cells
                g0 = -math.log(PML_R0) * math.log(PML_GRADE) / (2.0 * dl * (PML_GRADE ** (W / dl) - 1.0))
                pl = np.where(Dl > 0, g0 * PML_GRADE ** (np.maximum(Dl, 0) / dl), 0.0)
                pm = np.where(Dm > 0, g0 * PML_GRADE ** (np.maximum(Dm, 0) / dl), 0.0)
                gl += pl; gm[:-1] += pm
            self.g_line.append(gl); self.g_mid.append(gm)
            self.pml_lo.append(lo_cells); self.pml_hi.append(hi_cells)
        self.has_pml = any(c > 0 for c in self.pml_lo + self.pml_hi)

    def pml_boxes(self):
        """non-overlapping boxes (global node index ranges, half open) covering every PML position"""
        nx, ny, nz = self.n
        def rng(a):
            lo = (0, self.pml_lo[a]) if self.pml_lo[a] else None
            hi = (self.n[a] - 1 - self.pml_hi[a], self.n[a]) if self.pml_hi[a] else None
            mid = (self.pml_lo[a] if lo else 0, self.n[a] - 1 - self.pml_hi[a] if hi else self.n[a])
            return lo, hi, mid
        (xl, xh, xm), (yl, yh, ym), (zl, zh, zm) = rng(0), rng(1), rng(2)
        # z-slabs take whole planes, y-slabs whole x-rows of the remaining planes (both are fused into the volume
        # kernels, csrc/b200fdtd.cu RowParams); x-slabs are the narrow rest for the separate pre/post kernel
        boxes = []
        for r in (zl, zh):
            if r: boxes.append(((0, nx), (0, ny), r))
        for r in (yl, yh):
            if r: boxes.append(((0, nx), r, zm))
        for r in (xl, xh):
            if r: boxes.append((r, ym, zm))
        return [b for b in boxes if all(q[1] > q[0] for q in b)]

    # ------------------------------------------------------------------ EC blocks
    def _t(self, a, shape_axis):
        """1-D numpy array -> float64 tensor broadcastable along `shape_axis` of a [z,y,x] block"""
        t = torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(self.dev)
        shp = [1, 1, 1]; shp[2 - shape_axis] = -1
        return t.reshape(shp)

    def _quad_sum(self, pad, comp, rng):
        """sum over the 4 quadrants around comp-edges of prop * quarter area; block [nk,nj,ni] float64"""
        nP, nPP = (comp + 1) % 3, (comp + 2) % 3
        out = None
        for a_hi in (0, 1):
            for b_hi in (0, 1):
                sl = [None, None, None]
                for ax in range(3):
                    lo, hi = rng[ax]
                    if ax == comp:
                        off = 1
                    elif ax == nP:
                        off = a_hi
                    else:
                        off = b_hi
                    zoff = -self.kc0 if ax == 2 else 0
                    sl[2 - ax] = slice(lo + off + zoff, hi + off + zoff)
                ha = (self.h_hi if a_hi else self.h_lo)[nP][slice(*rng[nP])]
                hb = (self.h_hi if b_hi else self.h_lo)[nPP][slice(*rng[nPP])]
                term = pad[tuple(sl)].to(torch.float64) * self._t(ha, nP) * self._t(hb, nPP)
                out = term if out is None else out + term
        return out

    def ec_block(self, rng):
        """EC quantities on the node-index box rng=((i0,i1),(j0,j1),(k0,k1)) (half open).
        Returns dict of lists over the 3 components of float64 tensors [nk,nj,ni]:
          C, G (edge capacitance/conductance), epsE (effective eps_r at the edge),
          invL (1/face inductance; 0 where the H component is not updated), epsH."""
        (i0, i1), (j0, j1), (k0, k1) = rng
        if k0 - 1 < self.kc0 - 1 or k1 - 1 > self.kc1:
            raise ValueError(f"z-range {k0}:{k1} needs cells outside this builder's range {self.kc0}:{self.kc1}")
        out = {"C": [], "G": [], "epsE": [], "invL": [], "epsH": []}
        for comp in range(3):
            nP, nPP = (comp + 1) % 3, (comp + 2) % 3
            qe = self._quad_sum(self.eps_pad, comp, rng)
            qk = self._quad_sum(self.kap_pad, comp, rng)
            inv_len = self._t(1.0 / self.len_p[comp][slice(*rng[comp])], comp)
            area = self._t((self.h_lo[nP] + self.h_hi[nP])[slice(*rng[nP])], nP) * \
                   self._t((self.h_lo[nPP] + self.h_hi[nPP])[slice(*rng[nPP])], nPP)
            out["C"].append(EPS0 * qe * inv_len)
            out["G"].append(qk * inv_len)
            out["epsE"].append(torch.clamp(qe / area, min=1.0e-30))
            # face inductance L = mu0 * A / l~ ; H at the top index of any axis is never updated
            dP = np.concatenate([self.d[nP], [0.0]])[slice(*rng[nP])]
            dPP = np.concatenate([self.d[nPP], [0.0]])[slice(*rng[nPP])]
            ld = self.len_d[comp][slice(*rng[comp])].copy()
            top = np.arange(*rng[comp]) == self.n[comp] - 1
            A = self._t(dP, nP) * self._t(dPP, nPP)
            num = self._t(np.where(top, 0.0, ld), comp)
            invL = torch.where(A > 0, num / (MUE0 * torch.clamp(A, min=1e-300)), torch.zeros_like(A * num))
            out["invL"].append(invL)
            # eps at the H position: mean of the two cells adjacent along comp
            sl_lo, sl_hi = [None] * 3, [None] * 3
            for ax in range(3):
                lo, hi = rng[ax]
                zoff = -self.kc0 if ax == 2 else 0
                if ax == comp:
                    sl_lo[2 - ax] = slice(lo + zoff, hi + zoff); sl_hi[2 - ax] = slice(lo + 1 + zoff, hi + 1 + zoff)
                else:
                    sl_lo[2 - ax] = sl_hi[2 - ax] = slice(lo + 1 + zoff, hi + 1 + zoff)
            e_lo, e_hi = self.eps_pad[tuple(sl_lo)].to(torch.float64), self.eps_pad[tuple(sl_hi)].to(torch.float64)
            cnt = (e_lo > 0).to(torch.float64) + (e_hi > 0).to(torch.float64)
            out["epsH"].append(torch.where(cnt > 0, (e_lo + e_hi) / torch.clamp(cnt, min=1.0), torch.ones_like(cnt)))
        # lumped conductances replace the material conductance on their edges
        for (comp, ii_, jj, kk, G) in self.lumped_G:
            sel = (ii_ >= i0) & (ii_ < i1) & (jj >= j0) & (jj < j1) & (kk >= k0) & (kk < k1)
            if sel.any():
                out["G"][comp][torch.from_numpy(kk[sel] - k0).to(self.dev), torch.from_numpy(jj[sel] - j0).to(self.dev),
                               torch.from_numpy(ii_[sel] - i0).to(self.dev)] = torch.from_numpy(G[sel]).to(self.dev)
        return out

    def _pec_mask(self, comp, rng):
        (i0, i1), (j0, j1), (k0, k1) = rng
        m = torch.zeros((k1 - k0, j1 - j0, i1 - i0), dtype=torch.bool, device=self.dev)
        for (c, ri, rj, rk, mask) in self.pec:
            if c != comp:
                continue
            a0, a1 = max(ri[0], i0), min(ri[1], i1)
            b0, b1 = max(rj[0], j0), min(rj[1], j1)
            c0, c1 = max(rk[0], k0), min(rk[1], k1)
            if a1 <= a0 or b1 <= b0 or c1 <= c0:
                continue
            dst = (slice(c0 - k0, c1 - k0), slice(b0 - j0, b1 - j0), slice(a0 - i0, a1 - i0))
            if mask is None:
                m[dst] = True
            else:
                sub = mask[c0 - rk[0]:c1 - rk[0], b0 - rj[0]:b1 - rj[0], a0 - ri[0]:a1 - ri[0]]
                m[dst] |= torch.from_numpy(np.ascontiguousarray(sub)).to(self.dev)
        # outer boundary: tangential E on the min/max planes is PEC (also behind PML; Mur overrides via its lists)
        for ax in range(3):
            if ax == comp:
                continue
            lo, hi = rng[ax]
            sl = [slice(None)] * 3
            if lo == 0:
                sl[2 - ax] = slice(0, 1); m[tuple(sl)] = True
            if hi == self.n[ax]:
                sl[2 - ax] = slice(hi - lo - 1, hi - lo); m[tuple(sl)] = True
        return m

    def _rates(self, comp, rng, ec, field_kind):
        """PML rates r_a = g_a(position) * c0 for a = comp, nP, nPP at the positions of E (0) or H (1) comp.
        The stretching of a coordinate is a property of the coordinate, not of the material it runs through: with the
        local phase velocity instead of c0 the stretch factor jumps at a material interface that enters the PML (a
        substrate under a microstrip line) and the layer goes unstable (tests/test_oracle_physics.py::
        test_microstrip_line_impedance_and_effective_permittivity blows up within 4000 steps).  In a vacuum-filled PML both
        choices are the same number: that is the case in the live reference scenes (…microstrip_3d.py, …multi_3d.py,
        …fixed.py: substrate and ground plane end inside the domain).  It is NOT the case in the legacy backend
        (solver_fdtd_openems.py:209-217: substrate and ground plane span the whole SimBox into PML_8); there openEMS would
        use kappa/eps_eff of the material inside the layer, so agreement with openEMS for that backend is not exact and
        rests on this project's own oracle, which shares this formulation.
        B200FDTD_PML_LOCAL_C=1 restores the material-dependent rates for comparison."""
        eps = ec["epsE"][comp] if field_kind == 0 else ec["epsH"][comp]
        c_loc = C0 / torch.sqrt(eps) if os.environ.get("B200FDTD_PML_LOCAL_C") else torch.full_like(eps, C0)
        r = {}
        for a in range(3):
            on_mid = (a == comp) if field_kind == 0 else (a != comp)
            g = (self.g_mid if on_mid else self.g_line)[a][slice(*rng[a])]
            r[a] = self._t(g, a) * c_loc if g.any() else None
        return r

    # ------------------------------------------------------------------ time step
    def estimate_timestep(self, chunk=32):
        """dt = 2 / sqrt(max over nodes of S_x+S_y+S_z), S_n = (1/C_n) * sum of 1/L over the 4 faces around the edge.
        Equals dx/(c sqrt(3)) on a uniform vacuum grid (the openEMS 'Rennings' estimate reduces to the same there)."""
        nx, ny, nz = self.n
        k_lo, k_hi = self.K0, self.K1                      # owned planes only: the global dt is the min over ranks
        worst = 0.0
        k = k_lo
        while k < k_hi:
            k1 = min(k + chunk, k_hi)
            kk0 = max(k - 1, self.kc0)                     # one extra plane below for the k-1 neighbours (only its 1/L is used)
            ec = self.ec_block(((0, nx), (0, ny), (kk0, k1)))
            off = k - kk0
            tot = None
            for comp in range(3):
                nP, nPP = (comp + 1) % 3, (comp + 2) % 3
                def shifted(t, ax):
                    z = torch.zeros_like(t)
                    if ax == 0: z[:, :, 1:] = t[:, :, :-1]
                    elif ax == 1: z[:, 1:, :] = t[:, :-1, :]
                    else: z[1:] = t[:-1]
                    return z
                s = ec["invL"][nPP] + shifted(ec["invL"][nPP], nP) + ec["invL"][nP] + shifted(ec["invL"][nP], nPP)
                C = ec["C"][comp]
                live = (C > 0) & ~self._pec_mask(comp, ((0, nx), (0, ny), (kk0, k1)))     # PEC edges are never updated
                S = torch.where(live, s / torch.clamp(C, min=1e-300), torch.zeros_like(s))
                tot = S if tot is None else tot + S
            worst = max(worst, float(tot[off:].max()))
            k = k1
        return 2.0 / math.sqrt(worst)

    # ------------------------------------------------------------------ coefficients
    def coefficients(self, k0, k1, px, dt, chunk=32, compress=True):
        """vv, vi, ii, iv as float32 [3][k1-k0+2][ny][px] (ghost planes and pad columns zero) for node planes [k0,k1).
        With compress=True the row compression tables of both passes are left in self.row_compression."""
        nx, ny, nz = self.n
        shape = (3, k1 - k0 + 2, ny, px)
        vv = torch.zeros(shape, dtype=torch.float32, device=self.dev)
        vi, ii, iv = torch.zeros_like(vv), torch.zeros_like(vv), torch.zeros_like(vv)
        fE = RowFactor(nx, px, shape[1], ny, self.dev) if compress else None
        fH = RowFactor(nx, px, shape[1], ny, self.dev) if compress else None

        def put(dst, comp, sl, arr64, fac, slot):
            dst[comp, sl, :, :nx] = fac.factor(arr64, slot, sl.start) if fac is not None else arr64.to(torch.float32)

        k = k0
        while k < k1:
            ke = min(k + chunk, k1)
            rng = ((0, nx), (0, ny), (k, ke))
            ec = self.ec_block(rng)
            dst = slice(k - k0 + 1, ke - k0 + 1)
            for comp in range(3):
                nP = (comp + 1) % 3
                C, G = ec["C"][comp], ec["G"][comp]
                ok = C > 0
                Cs = torch.clamp(C, min=1e-300)
                x = dt * G / (2.0 * Cs)
                rE = self._rates(comp, rng, ec, 0)
                if rE[nP] is not None:
                    x = x + 0.5 * dt * rE[nP]
                pec = self._pec_mask(comp, rng)
                ok = ok & ~pec
                z = torch.zeros_like(C)
                put(vv, comp, dst, torch.where(ok, (1.0 - x) / (1.0 + x), z), fE, comp)
                put(vi, comp, dst, torch.where(ok, (dt / Cs) / (1.0 + x), z), fE, 3 + comp)
                invL = ec["invL"][comp]
                okh = invL > 0
                y = torch.zeros_like(invL)
                rH = self._rates(comp, rng, ec, 1)
                if rH[nP] is not None:
                    y = y + 0.5 * dt * rH[nP]
                put(ii, comp, dst, torch.where(okh, (1.0 - y) / (1.0 + y), z), fH, comp)
                put(iv, comp, dst, torch.where(okh, dt * invL / (1.0 + y), z), fH, 3 + comp)
            k = ke
        self.row_compression = None
        if compress:
            self.row_compression = {0: fE.tables(), 1: fH.tables()}
            self.row_compression_stats = {0: (fE.rows_compressed, fE.rows_total), 1: (fH.rows_compressed, fH.rows_total)}
        return vv, vi, ii, iv

    def pml_coefficients(self, box, dt, pad_to=None, compress=False):
        """second-stage UPML coefficients of one box ((i0,i1),(j0,j1),(k0,k1)); float32 [3][bz][by][bx] each, bx padded
        with zero columns to pad_to.  With compress=True the rows are factorised like the operator rows (RowFactor,
        9 slots: a_xyz, fo_xyz, fn_xyz) and the tables are returned as 'cmp_v' / 'cmp_i' = (xvecs, records)."""
        ec = self.ec_block(box)
        bx = box[0][1] - box[0][0]; by = box[1][1] - box[1][0]; bz = box[2][1] - box[2][0]
        bxp = int(pad_to) if pad_to else bx
        names = ("vv", "vvfo", "vvfn", "ii", "iifo", "iifn")
        out = {n: torch.zeros((3, bz, by, bxp), dtype=torch.float32, device=self.dev) for n in names}
        fac = [RowFactor(bx, bxp, bz, by, self.dev, nslots=9, rec_bytes=48) if compress else None for _ in range(2)]
        for comp in range(3):
            nPP = (comp + 2) % 3
            for kind, trip in ((0, ("vv", "vvfo", "vvfn")), (1, ("ii", "iifo", "iifn"))):
                r = self._rates(comp, box, ec, kind)
                base = ec["C"][comp] if kind == 0 else ec["invL"][comp]
                zero = torch.zeros_like(base)
                rn = r[comp] if r[comp] is not None else zero
                rpp = r[nPP] if r[nPP] is not None else zero
                den = 2.0 + dt * rpp
                vals = ((2.0 - dt * rpp) / den, (2.0 - dt * rn) / den, (2.0 + dt * rn) / den)
                for q, (name, v64) in enumerate(zip(trip, vals)):
                    v64 = v64 + zero                                   # broadcast to the full block
                    out[name][comp, :, :, :bx] = fac[kind].factor(v64, 3 * q + comp, 0) if compress else v64.to(torch.float32)
        if compress:
            out["cmp_v"], out["cmp_i"] = fac[0].tables(), fac[1].tables()
        return out

    # ------------------------------------------------------------------ narrow-band lists (global indices)
    def excitation_list(self, dt):
        """(comp, i, j, k, amp, delay_steps) arrays for soft E excitation boxes (App. A5)"""
        comp_l, i_l, j_l, k_l, amp_l, del_l = [], [], [], [], [], []
        tol = self._tol()
        for ex in self.s.excitations:
            lo_i = [_snap(self.x[a], ex["lo"][a]) for a in range(3)]
            hi_i = [_snap(self.x[a], ex["hi"][a]) for a in range(3)]
            for comp in range(3):
                v = ex["vec"][comp]
                if v == 0:
                    continue
                rng = [np.arange(lo_i[a], hi_i[a] + 1) for a in range(3)]
                rng[comp] = np.arange(lo_i[comp], hi_i[comp])
                if any(len(r) == 0 for r in rng):
                    continue
                I = np.meshgrid(*rng, indexing="ij")
                idx = [g.ravel() for g in I]
                comp_l.append(np.full(len(idx[0]), comp)); i_l.append(idx[0]); j_l.append(idx[1]); k_l.append(idx[2])
                amp_l.append(v * self.len_p[comp][idx[comp]])
                del_l.append(np.full(len(idx[0]), int(round(ex.get("delay", 0.0) / dt))))
        if not comp_l:
            z = np.zeros(0, np.int64)
            return z, z, z, z, np.zeros(0), z
        cat = np.concatenate
        return cat(comp_l), cat(i_l), cat(j_l), cat(k_l), cat(amp_l), cat(del_l)

    def mur_list(self, dt):
        """(comp, dst ijk, src ijk, coeff) for every Mur face; later faces win on shared edges (App. A3)"""
        out = []
        nx, ny, nz = self.n
        for ax in range(3):
            for side in (0, 1):
                if self.s.bc[2 * ax + side] != BC_MUR:
                    continue
                b = 0 if side == 0 else self.n[ax] - 1
                s = 1 if side == 0 else self.n[ax] - 2
                delta = abs(self.x[ax][b] - self.x[ax][s])
                if ax == 2 and not (self.K0 <= b < self.K1):
                    continue
                if ax == 2 and not (self.K0 <= s < self.K1):
                    raise ValueError("a z-slab holding a Mur face needs at least 2 planes")
                for comp in ((ax + 1) % 3, (ax + 2) % 3):
                    rng = [np.arange(self.n[a]) for a in range(2)] + [np.arange(self.K0, self.K1)]
                    rng[ax] = np.array([b])
                    I = np.meshgrid(*rng, indexing="ij")
                    dst = [g.ravel() for g in I]
                    src = [g.copy() for g in dst]; src[ax] = np.full_like(dst[ax], s)
                    # local phase velocity from the effective eps of the inward edge
                    r = [(0, self.n[0]), (0, self.n[1]), (self.K0, self.K1)]; r[ax] = (s, s + 1)
                    eps = self.ec_block(tuple(r))["epsE"][comp].cpu().numpy()          # [nk,nj,ni]
                    eps = np.transpose(eps, (2, 1, 0)).ravel()                         # ij-order like meshgrid
                    c = C0 / np.sqrt(np.maximum(eps, 1e-30))
                    coeff = (c * dt - delta) / (c * dt + delta)
                    out.append((np.full(len(dst[0]), comp), dst, src, coeff, np.full(len(dst[0]), 2 * ax + side)))
        if not out:
            return None
        comp = np.concatenate([o[0] for o in out])
        dst = [np.concatenate([o[1][a] for o in out]) for a in range(3)]
        src = [np.concatenate([o[2][a] for o in out]) for a in range(3)]
        coeff = np.concatenate([o[3] for o in out])
        self.mur_face = np.concatenate([o[4] for o in out])
        key = ((comp * nz + dst[2]) * ny + dst[1]) * nx + dst[0]
        _, first_rev = np.unique(key[::-1], return_index=True)
        keep = np.sort(len(key) - 1 - first_rev)
        self.mur_face = self.mur_face[keep]                # which boundary face an entry belongs to (for ordering only)
        return comp[keep], [d[keep] for d in dst], [s[keep] for s in src], coeff[keep]

    def probe_lists(self):
        """per probe: (name, kind, comp[], i[], j[], k[], w[]) — voltage line / current loop integrals (App. A5)"""
        res = []
        for pr in self.s.probes:
            if pr["p_type"] == 0:
                start = [_snap(self.x[a], pr["start"][a]) for a in range(3)]
                stop = [_snap(self.x[a], pr["stop"][a]) for a in range(3)]
                comp, I, J, K, W = [], [], [], [], []
                pos = list(start)
                for n in range(3):                       # walk x, then y, then z (CalcVoltageIntegral)
                    if start[n] < stop[n]:
                        for q in range(start[n], stop[n]):
                            p = list(pos); p[n] = q
                            comp.append(n); I.append(p[0]); J.append(p[1]); K.append(p[2]); W.append(pr["weight"])
                    elif start[n] > stop[n]:
                        for q in range(stop[n], start[n]):
                            p = list(pos); p[n] = q
                            comp.append(n); I.append(p[0]); J.append(p[1]); K.append(p[2]); W.append(-pr["weight"])
                    pos[n] = stop[n]
                res.append((pr["name"], 0, np.array(comp), np.array(I), np.array(J), np.array(K), np.array(W, np.float64)))
            elif pr["p_type"] == 1:
                nd = pr["norm_dir"]
                if nd not in (0, 1, 2):
                    raise ValueError("current probe needs norm_dir")
                nP, nPP = (nd + 1) % 3, (nd + 2) % 3
                lo = np.minimum(pr["start"], pr["stop"]); hi = np.maximum(pr["start"], pr["stop"])
                tol = self._tol()
                st, sp = [0, 0, 0], [0, 0, 0]
                st[nd] = sp[nd] = int(np.argmin(np.abs(self.mid[nd] - 0.5 * (lo[nd] + hi[nd]))))
                for a in (nP, nPP):                       # smallest dual rectangle enclosing the box
                    below = np.where(self.mid[a] < lo[a] - tol)[0]
                    above = np.where(self.mid[a] > hi[a] + tol)[0]
                    if len(below) == 0 or len(above) == 0:
                        raise ValueError("current probe box touches the mesh boundary")
                    st[a], sp[a] = int(below[-1]), int(above[0])
                comp, I, J, K, W = [], [], [], [], []
                def add(c, p, w):
                    comp.append(c); I.append(p[0]); J.append(p[1]); K.append(p[2]); W.append(w * pr["weight"])
                # CalcCurrentIntegral: nP-directed dual edges at nPP = start (+) and stop (-);
                # nPP-directed dual edges at nP = start (-) and stop (+)
                for q in range(st[nP] + 1, sp[nP] + 1):
                    p = [0, 0, 0]; p[nd] = st[nd]; p[nP] = q
                    p[nPP] = st[nPP]; add(nP, p, +1.0)
                    p = list(p); p[nPP] = sp[nPP]; add(nP, p, -1.0)
                for q in range(st[nPP] + 1, sp[nPP] + 1):
                    p = [0, 0, 0]; p[nd] = st[nd]; p[nPP] = q
                    p[nP] = st[nP]; add(nPP, p, -1.0)
                    p = list(p); p[nP] = sp[nP]; add(nPP, p, +1.0)
                res.append((pr["name"], 1, np.array(comp), np.array(I), np.array(J), np.array(K), np.array(W, np.float64)))
            else:
                raise NotImplementedError("only voltage (0) and current (1) probes are used by the reference's ports")
        return res

    def nf2ff_faces(self):
        """6 faces of the Huygens box in global node indices (App. A6)"""
        nf = self.s.nf2ff
        if nf is None:
            return []
        lo = [_snap(self.x[a], nf["start"][a]) for a in range(3)]
        hi = [_snap(self.x[a], nf["stop"][a]) for a in range(3)]
        faces = []
        # a face that lies on a PEC / PMC boundary is dropped and the far field takes the image of the remaining faces in
        # that wall instead (openEMS nf2ff 'mirror', App. A6): self.nf2ff_mirrors = [(axis, wall coordinate, bc type)]
        self.nf2ff_mirrors = []
        for n in range(3):
            a, b = (n + 1) % 3, (n + 2) % 3
            for side, plane in ((0, lo[n]), (1, hi[n])):
                bc = self.s.bc[2 * n + side]
                on_wall = plane == (0 if side == 0 else self.n[n] - 1)
                if bc in (BC_PEC, BC_PMC) and on_wall:
                    self.nf2ff_mirrors.append((n, float(self.x[n][plane]), int(bc)))
                    continue
                faces.append(dict(normal=n, side=side, plane=plane, coord=float(self.x[n][plane]), a0=lo[a], a1=hi[a], b0=lo[b], b1=hi[b]))
        return faces

    def face_weights(self, face):
        """node positions [m] and trapezoid area weights of one face: (xa[na], xb[nb], wa[na], wb[nb])"""
        a, b = (face["normal"] + 1) % 3, (face["normal"] + 2) % 3
        def w(ax, i0, i1):
            x = self.x[ax][i0:i1 + 1]
            d = np.diff(x)
            return x, 0.5 * (np.concatenate([[0.0], d]) + np.concatenate([d, [0.0]]))
        xa, wa = w(a, face["a0"], face["a1"])
        xb, wb = w(b, face["b0"], face["b1"])
        return xa, xb, wa, wb

    def nyquist(self, dt):
        fmax = self.s.f0 + self.s.fc
        if fmax <= 0:
            return 100
        return max(1, int(math.floor(1.0 / (2.0 * fmax * dt))))
