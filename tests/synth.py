"""Seeded synthetic engine-level problems shared by the oracle and the CUDA engine (test helper)."""
import numpy as np


def lin(nz, ny, px, c, k, j, i):
    return ((c * (nz + 2) + (k + 1)) * ny + j) * px + i


def make_problem(nx=37, ny=29, nz=23, px=40, seed=0, with_pml=True, with_mur=True, with_exc=True,
                 with_probes=True, with_nf2ff=True, interval=3, nfreq=3, fused_pml=False, compress=False):
    rng = np.random.default_rng(seed)
    shape = (3, nz + 2, ny, px)
    P = {"nx": nx, "ny": ny, "nz": nz, "px": px, "shape": shape, "interval": interval}

    def coef(lo, hi):
        a = rng.uniform(lo, hi, shape).astype(np.float32)
        a[..., nx:] = 0.0                                 # pad columns
        return a
    vv, ii = coef(0.9, 1.0), coef(0.9, 1.0)
    vi, iv = coef(0.05, 0.25), coef(0.05, 0.25)
    pec = rng.random(shape) < 0.05                        # random PEC edges
    vv[pec] = 0; vi[pec] = 0
    # H components on the top index planes do not exist
    for a in (ii, iv):
        a[:, :, ny - 1, :] = 0; a[:, :, :, nx - 1:] = 0
    if compress:
        # row compression (b200fdtd_set_row_compression): ~70 % of the rows become scale*xvec exactly; a few records
        # make false claims and must be demoted by the device-side verification
        P["cmp"] = {}
        for which, (ca, cb) in enumerate(((vv, vi), (ii, iv))):
            nvec = 5
            xv = np.zeros((nvec, px), np.float32)
            xv[:, :nx] = rng.uniform(0.5, 1.0, (nvec, nx)).astype(np.float32)
            xv[0, :nx] = 1.0
            if which == 1:
                xv[:, nx - 1:] = 0.0
            meta_f = np.zeros((nz + 2, ny, 8), np.float32)
            meta = meta_f.view(np.uint8)
            meta[:, :, 24:30] = 255
            wrong = 0
            for slot in range(6):
                arr = (ca, cb)[slot // 3][slot % 3]
                for k in range(nz):
                    for j in range(ny):
                        u = rng.random()
                        if u < 0.7:
                            vid = int(rng.integers(0, nvec)); scv = np.float32(rng.uniform(0.05, 1.0))
                            arr[k + 1, j, :] = scv * xv[vid]          # float32 product
                            meta_f[k + 1, j, slot] = scv; meta[k + 1, j, 24 + slot] = vid
                        elif u < 0.73:
                            vid = int(rng.integers(0, nvec)); scv = np.float32(rng.uniform(0.05, 1.0))
                            meta_f[k + 1, j, slot] = scv; meta[k + 1, j, 24 + slot] = vid   # false claim: row left as is
                            wrong += 1
            P["cmp"][which] = dict(xvecs=xv, meta=meta.copy(), wrong=wrong)
        pec = pec & (rng.random(shape) < 0.0)              # keep compressed rows intact (no random PEC on top)
        for a in (ii, iv):
            a[:, :, ny - 1, :] = 0; a[:, :, :, nx - 1:] = 0
        # a false claim may have become true by the zeroing above only if the whole row is zero: recount exactly
        for which, (ca, cb) in enumerate(((vv, vi), (ii, iv))):
            c = P["cmp"][which]; mf = c["meta"].view(np.float32); wrong = 0; good = 0
            for slot in range(6):
                arr = (ca, cb)[slot // 3][slot % 3]
                ids = c["meta"][:, :, 24 + slot]
                for k in range(nz):
                    for j in range(ny):
                        vid = ids[k + 1, j]
                        if vid == 255:
                            continue
                        ok = np.array_equal((mf[k + 1, j, slot] * c["xvecs"][vid]).view(np.uint32), arr[k + 1, j].view(np.uint32))
                        wrong += (not ok); good += ok
            c["wrong"], c["good"] = wrong, good
    P.update(vv=vv, vi=vi, ii=ii, iv=iv)
    volt = rng.standard_normal(shape).astype(np.float32); volt[..., nx:] = 0
    curr = rng.standard_normal(shape).astype(np.float32); curr[..., nx:] = 0
    volt[vv == 0] *= (vv[vv == 0] != 0)                   # keep PEC edges at zero initially
    P.update(volt0=volt, curr0=curr)
    total = 3 * (nz + 2) * ny * px

    def rand_idx(n):
        c = rng.integers(0, 3, n); k = rng.integers(0, nz, n); j = rng.integers(0, ny, n); i = rng.integers(0, nx, n)
        return lin(nz, ny, px, c, k, j, i).astype(np.int64)

    if with_exc:
        n = 12
        idx = np.unique(rand_idx(n))
        P["exc"] = dict(idx=idx, amp=rng.uniform(-1, 1, len(idx)).astype(np.float32),
                        delay=rng.integers(0, 4, len(idx)).astype(np.int32),
                        signal=np.concatenate([[0.0], rng.standard_normal(15)]).astype(np.float32))
    if with_mur:
        dst, src = [], []
        for axis in range(3):
            n_ax = (nx, ny, nz)[axis]
            for b, s in ((0, 1), (n_ax - 1, n_ax - 2)):
                for comp in ((axis + 1) % 3, (axis + 2) % 3):
                    rng_ax = [np.arange(nx), np.arange(ny), np.arange(nz)]
                    rng_ax[axis] = np.array([b])
                    I, J, K = np.meshgrid(*rng_ax, indexing="ij")
                    d = lin(nz, ny, px, comp, K, J, I).ravel()
                    idx3 = [I, J, K]; idx3[axis] = np.full_like(I, s)
                    sidx = lin(nz, ny, px, comp, idx3[2], idx3[1], idx3[0]).ravel()
                    dst.append(d); src.append(sidx)
        dst = np.concatenate(dst); src = np.concatenate(src)
        # later faces override earlier ones on shared edges: keep the last entry per dst
        _, first_rev = np.unique(dst[::-1], return_index=True)
        keep = np.sort(len(dst) - 1 - first_rev)
        dst, src = dst[keep], src[keep]
        P["mur"] = dict(dst=dst.astype(np.int64), src=src.astype(np.int64),
                        coeff=rng.uniform(-0.5, 0.5, len(dst)).astype(np.float32))
    if with_pml:
        boxes = []
        if fused_pml:
            # z-slabs over whole planes, y-slabs over whole rows of the planes in between (fused into the volume kernels),
            # plus narrow x-slabs for the separate pass
            layout = ((0, 0, 0, px, ny, 3), (0, 0, nz - 4, px, ny, 4), (0, 0, 3, px, 4, nz - 7), (0, ny - 5, 3, px, 5, nz - 7),
                      (0, 4, 3, 8, ny - 9, nz - 7), ((nx - 6) // 4 * 4, 4, 3, (nx - (nx - 6) // 4 * 4 + 3) // 4 * 4, ny - 9, nz - 7))
        else:
            layout = ((0, 0, 0, 5, ny, nz), (nx - 6, 0, 0, 6, ny, nz), (5, 0, 0, nx - 11, 4, nz), (5, 4, nz - 5, nx - 11, ny - 4, 5))
        for (x0, y0, z0, bx, by, bz) in layout:
            shp = (3, bz, by, bx)
            B = dict(x0=x0, y0=y0, z0=z0, bx=bx, by=by, bz=bz)
            for name in ("vv", "vvfo", "vvfn", "ii", "iifo", "iifn"):
                B[name] = rng.uniform(0.6, 1.0, shp).astype(np.float32)
                if x0 + bx > nx:
                    B[name][..., nx - x0:] = 0.0              # pad columns of full-row slabs
            boxes.append(B)
        P["pml"] = boxes
    if with_probes:
        kinds, offs, idxs, ws = [], [0], [], []
        for p, n in enumerate((4, 9, 300)):
            kinds.append(p % 2)
            idxs.append(rand_idx(n)); ws.append(rng.choice([-1.0, 1.0], n).astype(np.float32))
            offs.append(offs[-1] + n)
        P["probes"] = dict(kind=np.array(kinds, np.int32), offset=np.array(offs, np.int64), idx=np.concatenate(idxs),
                           weight=np.concatenate(ws), freqs=np.linspace(1e9, 3e9, nfreq), dt=1.1e-12, max_samples=64)
    if with_nf2ff:
        faces = []
        lo = (3, 3, 2); hi = (nx - 4, ny - 4, nz - 3)
        for n in range(3):
            a, b = (n + 1) % 3, (n + 2) % 3
            for plane in (lo[n], hi[n]):
                faces.append(dict(normal=n, plane=plane, a0=lo[a], a1=hi[a], b0=lo[b], b1=hi[b]))
        il = [rng.uniform(500, 2000, m).astype(np.float32) for m in (nx, ny, nz + 2)]
        idl = [rng.uniform(500, 2000, m).astype(np.float32) for m in (nx, ny, nz + 2)]
        P["nf2ff"] = dict(faces=faces, freqs=np.array([2.45e9, 3.1e9][:max(1, min(2, nfreq))]), dt=1.1e-12, inv_len=il, inv_dual=idl)
    return P


def apply(E, P, np_fields=True):
    """Load problem P into an engine E (oracle RefEngine or CUDA Engine; same method names)."""
    E.set_coeffs(P["vv"], P["vi"], P["ii"], P["iv"])
    if "cmp" in P and hasattr(E, "set_row_compression"):
        E.cmp_counts = {w: E.set_row_compression(w, c["xvecs"], c["meta"].copy()) for w, c in P["cmp"].items()}
    if np_fields:
        E.volt[...] = P["volt0"]; E.curr[...] = P["curr0"]
    else:
        import torch
        E.volt.copy_(torch.from_numpy(P["volt0"])); E.curr.copy_(torch.from_numpy(P["curr0"]))
    if "exc" in P:
        E.set_excitation(**P["exc"])
    if "mur" in P:
        E.set_mur(**P["mur"])
    if "pml" in P:
        E.set_pml(P["pml"])
    if "probes" in P:
        pr = P["probes"]
        E.set_probes(pr["kind"], pr["offset"], pr["idx"], pr["weight"], P["interval"], pr["max_samples"], pr["freqs"], pr["dt"])
    if "nf2ff" in P:
        nf = P["nf2ff"]
        E.set_nf2ff(nf["faces"], nf["freqs"], P["interval"], nf["dt"], nf["inv_len"], nf["inv_dual"])
