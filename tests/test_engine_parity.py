"""CUDA engine vs CPU oracle on identical seeded engine-level inputs (through the C-ABI).

Bar: volt/curr bit-exact (fp32, same operation order); probe series and DFT accumulators within
the stated tolerance of the oracle's double-precision accumulators.
"""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def _engines(P):
    from oracle.fdtd_ref import RefEngine
    from b200fdtd.engine import Engine
    R = RefEngine(P["nx"], P["ny"], P["nz"], P["px"])
    G = Engine(P["nx"], P["ny"], P["nz"], P["px"])
    synth.apply(R, P, True)
    synth.apply(G, P, False)
    return R, G


def _assert_fields_equal(R, G, what=""):
    gv, gc = G.volt.cpu().numpy(), G.curr.cpu().numpy()
    nz = R.nz
    # owned planes must agree bit for bit; ghost planes must be untouched
    assert np.array_equal(gv[:, 1:nz + 1].view(np.uint32), R.volt[:, 1:nz + 1].view(np.uint32)), "volt differs " + what
    assert np.array_equal(gc[:, 1:nz + 1].view(np.uint32), R.curr[:, 1:nz + 1].view(np.uint32)), "curr differs " + what
    assert np.array_equal(gv[:, [0, nz + 1]], R.volt[:, [0, nz + 1]])
    assert np.array_equal(gc[:, [0, nz + 1]], R.curr[:, [0, nz + 1]])


@pytest.mark.parametrize("shape", [(37, 29, 23, 40), (128, 16, 9, 128), (130, 9, 5, 160), (5, 4, 3, 8), (257, 33, 17, 288)])
@pytest.mark.parametrize("tune", [(16, 4), (1, 8), (5, 2), (64, 1), (3, 16)])
def test_volume_kernels_bit_exact(shape, tune):
    nx, ny, nz, px = shape
    P = synth.make_problem(nx, ny, nz, px, seed=nx + ny, with_pml=False, with_mur=False, with_exc=False,
                           with_probes=False, with_nf2ff=False)
    R, G = _engines(P)
    G.set_tuning(kz=tune[0], ty=tune[1])
    for it in range(3):
        R.update_only(0); G.update_only(0)
        _assert_fields_equal(R, G, f"after E update {it}")
        R.update_only(1); G.update_only(1)
        _assert_fields_equal(R, G, f"after H update {it}")


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("fused,variant", [(False, 0), (True, 0), (True, 1), (True, 2), (True, 8)])
def test_full_step_all_extensions(use_graph, fused, variant):
    """fused=True lays the PML out as whole-row slabs: variant 0 folds them into the volume kernels,
    variant 1 forces the separate pre/post passes, variant 2 keeps the fused launches on the main stream
    (no side-stream fork), variant 8 leaves the narrow x-slabs to the separate kernel (no x-edge launches);
    all must equal the oracle bit for bit"""
    P = synth.make_problem(37, 29, 23, 40, seed=3, fused_pml=fused)
    R, G = _engines(P)
    G.set_tuning(kz=5, ty=4, variant=variant)
    n = 20
    R.run(n); G.run(n, use_graph=use_graph)
    _assert_fields_equal(R, G, "after 20 full steps")
    assert G.ts == R.ts == n
    ns = n // P["interval"]
    assert G.num_samples == ns
    s_g = G.series.cpu().numpy()[:, :ns].astype(np.float64)
    s_r = R.series[:, :ns]
    scale = np.abs(s_r).max()
    assert np.abs(s_g - s_r).max() <= 2e-6 * scale * 300 ** 0.5, "probe series"
    d_g = G.probe_dft.cpu().numpy().astype(np.float64)
    assert np.abs(d_g - R.probe_dft).max() <= 1e-5 * np.abs(R.probe_dft).max(), "probe DFT"
    for fa_g, fa_r in zip(G.face_acc, R.face_acc):
        a = fa_g.cpu().numpy().astype(np.float64)
        assert np.abs(a - fa_r).max() <= 1e-5 * np.abs(fa_r).max(), "NF2FF face DFT"
    e_r, e_g = R.energy(), G.energy()
    assert abs(e_g - e_r) <= 1e-6 * abs(e_r)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("tune", [(16, 4, 0), (3, 8, 0), (5, 2, 4)])
def test_row_compression_is_lossless(fused, tune):
    """compressed rows (scale*xvec) give bit-identical fields; false claims are demoted by the device-side check;
    variant 4 ignores the tables"""
    P = synth.make_problem(37, 29, 23, 40, seed=9, fused_pml=fused, compress=True)
    R, G = _engines(P)
    for w in (0, 1):
        good, demoted = G.cmp_counts[w]
        assert demoted == P["cmp"][w]["wrong"] and good == P["cmp"][w]["good"], (w, G.cmp_counts[w], P["cmp"][w]["wrong"])
        assert good > 1000
    G.set_tuning(kz=tune[0], ty=tune[1], variant=tune[2])
    if not fused:          # update_only on fused slabs includes their PML pre/post, which the bare oracle update does not
        for it in range(2):
            R.update_only(0); G.update_only(0)
            _assert_fields_equal(R, G, f"after E update {it}")
            R.update_only(1); G.update_only(1)
            _assert_fields_equal(R, G, f"after H update {it}")
    R.run(11); G.run(11, use_graph=True)
    _assert_fields_equal(R, G, "after 11 full steps with row compression")


def test_half_steps_equal_run():
    P = synth.make_problem(33, 17, 11, 64, seed=5)
    R, G = _engines(P)
    for _ in range(7):
        G.half_step(0); G.half_step(1); G.half_step(2)
    R.run(7)
    _assert_fields_equal(R, G, "half steps")


def test_graph_resume_and_misaligned_chunks():
    P = synth.make_problem(21, 13, 9, 32, seed=7, interval=4)
    R, G = _engines(P)
    for n in (1, 2, 9, 4, 3):          # crosses chunk boundaries at odd offsets
        G.run(n, use_graph=True); R.run(n)
        _assert_fields_equal(R, G, f"after +{n}")
    ns = R.ts // 4
    assert np.abs(G.series.cpu().numpy()[:, :ns] - R.series[:, :ns]).max() <= 1e-4 * np.abs(R.series).max()


def test_farfield_kernel():
    from oracle import fdtd_ref
    from b200fdtd.engine import farfield
    rng = np.random.default_rng(0)
    n = 5000
    pos = rng.uniform(-0.1, 0.1, (3, n))
    J = rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))
    M = rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))
    th = np.deg2rad(np.arange(0, 181, 10.0)); ph = np.deg2rad(np.full_like(th, 35.0))
    k = 2 * np.pi * 2.45e9 / 299792458.0
    ref = fdtd_ref.farfield(pos, J, M, k, th, ph)
    got = farfield(pos, J, M, k, th, ph)
    for r, g in zip(ref, got):
        assert np.abs(r - g).max() <= 2e-5 * np.abs(np.concatenate(ref)).max()


def test_errors_are_reported_not_fatal():
    from b200fdtd.engine import Engine
    from b200fdtd import B200FDTDError
    with pytest.raises(B200FDTDError):
        Engine(1, 4, 4, 4)                       # grid too small
    G = Engine(8, 8, 4, 8)
    with pytest.raises(B200FDTDError):
        G.run(1)                                 # coefficients not bound
    with pytest.raises(B200FDTDError):
        G.set_excitation([10 ** 12], [1.0], [0], [0.0, 1.0])   # index out of range


def test_degenerate_inputs():
    """no extensions at all, zero steps, a single plane, pitch wider than the grid"""
    from oracle.fdtd_ref import RefEngine
    from b200fdtd.engine import Engine
    for (nx, ny, nz, px) in ((9, 3, 1, 12), (4, 2, 2, 32), (33, 5, 2, 64)):
        P = synth.make_problem(nx, ny, nz, px, seed=1, with_pml=False, with_mur=False, with_exc=False, with_probes=False, with_nf2ff=False)
        R, G = _engines(P)
        G.run(0); R.run(0)
        _assert_fields_equal(R, G, "zero steps")
        G.run(5, use_graph=True); R.run(5)
        _assert_fields_equal(R, G, f"bare grid {nx}x{ny}x{nz}")
        assert G.ts == 5 and G.num_samples == 0
    # empty lists are accepted and switch the feature off
    G = Engine(8, 8, 4, 8)
    G.set_excitation([], [], [], [0.0])
    G.set_mur([], [], [])
    G.set_pml([])


@pytest.mark.parametrize("shape", [(37, 29, 23, 40), (261, 21, 14, 288), (130, 12, 9, 160)])
@pytest.mark.parametrize("layout", ["slabs", "mur", "bare"])
@pytest.mark.parametrize("interval,tile", [(2, (7, 32)), (3, (3, 2)), (4, (15, 5)), (7, (7, 1))])
def test_fused_h_to_e_launches_bit_exact(shape, layout, interval, tile):
    """graph chunks run H(n)+E(n+1) in one sweep over the plain region (update_he_kernel, second field copy): fields must
    equal the oracle's separate passes bit for bit for even and odd numbers of fused launches per chunk, several
    x-segments (nx > 124), PML slabs on every side (halo cells read from the slab launches' output), Mur and excitation
    between the steps, and every tile shape"""
    nx, ny, nz, px = shape
    kw = dict(with_pml=layout == "slabs", fused_pml=layout == "slabs", with_mur=layout != "bare", with_exc=layout != "bare",
              with_probes=True, with_nf2ff=False, interval=interval)
    P = synth.make_problem(nx, ny, nz, px, seed=nx + interval, **kw)
    R, G = _engines(P)
    G.set_he_tuning(*tile)
    n = 2 * interval + 1                   # two graph chunks and one eager step
    R.run(n); G.run(n, use_graph=True)
    assert G.he_active
    _assert_fields_equal(R, G, f"{layout} interval {interval} tile {tile}")
    G.set_tuning(variant=128)              # same again without the fused launches
    R.run(2 * interval - 1); G.run(2 * interval - 1, use_graph=True)     # eager up to the chunk boundary, then one graph chunk
    assert not G.he_active
    _assert_fields_equal(R, G, "unfused chunk after fused chunks")


@pytest.mark.parametrize("shape", [(37, 29, 23, 40), (261, 21, 14, 288), (130, 12, 9, 160)])
@pytest.mark.parametrize("layout", ["slabs", "mur"])
@pytest.mark.parametrize("interval,tile,variant,de", [(4, (7, 32), 0, 0), (3, (3, 2), 0, 0), (5, (15, 5), 0, 0), (2, (3, 1), 0, 0),
                                                      (4, (7, 32), 1 << 23, 1), (3, (7, 4), 1 << 23, 2), (5, (7, 3), 1 << 22, 0), (3, (15, 2), 1 << 22, 1),
                                                      (4, (7, 32), 512, 0), (5, (7, 32), 1 << 20, 0), (7, (7, 1), 0, 2)])
def test_fused_h_to_e_with_row_compression(shape, layout, interval, tile, variant, de):
    """the fused launch on a row-compressed operator (update_he6_kernel: planes staged by TMA bulk copies, whole-row PML slabs
    swept inside the launch with a double-buffered current flux; de = E planes landing ahead, 0 = automatic; bit 22 keeps the
    slabs as separate launches, 512 = the plain fusion update_he_kernel, bit 20 the overlapped slab schedule) on one and
    several x-segments, with rows whose compression claims are false (demoted on the device), even and odd numbers of H
    passes per chunk (the flux copies swap every pass)"""
    nx, ny, nz, px = shape
    P = synth.make_problem(nx, ny, nz, px, seed=9 + nx, with_pml=layout == "slabs", fused_pml=layout == "slabs",
                           compress=True, interval=interval, with_nf2ff=False)
    R, G = _engines(P)
    G.set_tuning(variant=variant)
    G.set_he_tuning(*tile, de=de)
    n = 2 * interval + 1
    R.run(n); G.run(n, use_graph=True)
    assert G.he_active
    _assert_fields_equal(R, G, f"fused H->E with row compression, variant {variant}")
    R.run(3); G.run(3, use_graph=False)          # eager spans continue from the same flux state
    _assert_fields_equal(R, G, "eager steps after the graph chunks")


@pytest.mark.parametrize("layout", ["slabs", "mur"])
@pytest.mark.parametrize("shape", [(37, 29, 23, 40), (261, 21, 14, 288)])
def test_fused_step_parts_of_a_z_slab_rank(layout, shape):
    """the z-slab form of the fused step (b200fdtd_fused_step_part: boundary planes by separate launches, interior planes by
    the fused launch, caller-owned second copy) on a single slab with zero ghost planes must equal the oracle's steps"""
    nx, ny, nz, px = shape
    P = synth.make_problem(nx, ny, nz, px, seed=5, with_pml=layout == "slabs", fused_pml=layout == "slabs",
                           with_probes=False, with_nf2ff=False)
    R, G = _engines(P)
    G.bind_alt_fields()
    # the synthetic problem has (static, non-zero) ghost planes: a real z-slab run receives them into whichever copy is
    # current, here they are simply mirrored into the second copy
    G.volt2[:, [0, nz + 1]] = G.volt[:, [0, nz + 1]]
    G.curr2[:, [0, nz + 1]] = G.curr[:, [0, nz + 1]]
    import torch
    torch.cuda.synchronize()
    n = 6
    G.half_step_part(0, 0); G.half_step_part(0, 1)            # E(0)
    for _ in range(n - 1):                                     # H(s-1) + E(s)
        for part in ((0, 1, 2, 3) if _ % 2 else (0, 4, 1, 2, 3)):     # with and without the split fused launch
            G.fused_step_part(part)
    G.half_step_part(1, 0); G.half_step_part(1, 1)            # H(n-1)
    G.sync()                                                   # the part calls do not join the engine stream with torch's
    vcur, ccur = G.current_copy()
    assert (vcur, ccur) == ((n - 1) % 2, (n - 1) % 2)
    if vcur:
        G.volt.copy_(G.volt2)
    if ccur:
        G.curr.copy_(G.curr2)
    G.reset_current_copy()
    R.run(n)
    assert G.ts == n
    _assert_fields_equal(R, G, f"{n} steps through the fused step parts")


def test_overlapping_pml_boxes_are_rejected():
    """every cell belongs to at most one PML box: the C-ABI refuses overlapping boxes instead of updating cells twice"""
    from b200fdtd.engine import Engine
    from b200fdtd import B200FDTDError
    P = synth.make_problem(37, 29, 23, 40, seed=3, fused_pml=True)
    G = Engine(P["nx"], P["ny"], P["nz"], P["px"])
    G.set_coeffs(P["vv"], P["vi"], P["ii"], P["iv"])
    boxes = [dict(b) for b in P["pml"]]
    boxes.append(dict(boxes[0]))                  # the first box once more
    with pytest.raises(B200FDTDError):
        G.set_pml(boxes)


def test_operator_built_on_the_gpu_equals_the_cpu_build():
    """the operator build runs as torch tensor code on the GPU for large slabs; the arrays it produces
    (coefficients, row-compression tables, PML slab coefficients) must be the ones the CPU build produces"""
    import torch
    import replay
    from b200fdtd.operator import OperatorBuilder
    F = replay.replay("trace_multi2_mur_q2")["FDTD"]          # rotated / translated boxes, thick copper, two ports
    S = F._setup()
    out = {}
    for dev in ("cpu", "cuda"):
        B = OperatorBuilder(S, device=torch.device(dev))
        dt = B.estimate_timestep()
        px = (B.n[0] + 31) // 32 * 32
        co = [t.cpu() for t in B.coefficients(0, B.n[2], px, dt)]
        tabs = [(None if xv is None else xv.cpu(), None if meta is None else meta.cpu()) for xv, meta in (B.row_compression[0], B.row_compression[1])]
        out[dev] = (dt, co, tabs)
    assert out["cpu"][0] == out["cuda"][0], "time step differs"
    for name, a, b in zip(("vv", "vi", "ii", "iv"), out["cpu"][1], out["cuda"][1]):
        diff = int((a.view(torch.int32) != b.view(torch.int32)).sum())
        assert diff == 0, f"{name}: {diff} coefficients differ between the GPU and the CPU build"
    for (xa, ma), (xb, mb) in zip(out["cpu"][2], out["cuda"][2]):
        assert torch.equal(xa, xb) and torch.equal(ma, mb)
