"""Profiling driver: a few time steps around one sampling point plus the post-processing kernels, between
cudaProfilerStart/Stop, so that `ncu --profile-from-start off` sees every kernel of the path exactly where it runs:
fused H->E launch, E/H volume launches, PML slab launches, Mur, excitation, probes, NF2FF running DFT, energy,
time-domain NF2FF DFT, far field.

  python tools/prof_step.py --workload patch100m|cube|config1|array16 [--cells N] [--steps 3]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "fdtd-solver-antennas_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="patch100m")
ap.add_argument("--cells", type=float, default=100e6)
ap.add_argument("--cube-n", type=int, default=768)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--seeded", action="store_true", help="per-cell seeded coefficients, row compression off")
args = ap.parse_args()

from b200fdtd import scenes  # noqa: E402

nf = port = None
if args.workload == "cube":
    F = scenes.vacuum_cube(args.cube_n, nrts=10 ** 5)
elif args.workload == "config1":
    import replay
    R = replay.replay("trace_single_pml8_q3")
    F, nf, port = R["FDTD"], R["nf"], R["FDTD"].ports[0]
elif args.workload == "array16":
    import replay
    R = replay.replay("trace_array16_mur_q1")
    F, nf, port = R["FDTD"], R["nf"], R["FDTD"].ports[0]
else:
    F, nf, port = scenes.patch_scene(target_cells=args.cells, boundary="PML_8", f0=2.5e9, fc=1.5e9, nrts=10 ** 5, end_criteria=1e-12,
                                     nf2ff_freqs=[2.45e9])
path = f"/tmp/b200fdtd_prof_{os.getpid()}"
F.Run(path, setup_only=True, cleanup=True)
sim = F._prepared[1]
E = sim.engine
if args.variant:
    E.set_tuning(variant=args.variant)
if args.seeded:
    g = torch.Generator(device=E.device).manual_seed(0)
    for arr, lo in ((E.vv, 0.9), (E.ii, 0.9), (E.vi, 0.5), (E.iv, 0.5)):
        for c in range(3):
            for k0 in range(0, arr.shape[1], 32):
                blk = arr[c, k0:k0 + 32]
                blk.mul_(torch.empty_like(blk).uniform_(lo, 1.0, generator=g))
    for which in (0, 1):
        E.set_row_compression(which, None, None)
iv = sim.interval
E.run(2 * iv, use_graph=True)                                   # warm: graph path, allocations, second field copy
E.run(iv - args.steps, use_graph=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
E.run(args.steps + 1, use_graph=False)                          # crosses the sampling point: probes + NF2FF DFT launch
E.update_only(2, join=False); E.update_only(3, join=False)      # the plain E and H launches on their own
e = sim.energy()
torch.cuda.synchronize()
if nf is not None:
    sim.collect()
    from openEMS import _registry
    res = sim.results
    res["device"] = 0; res["farfield_fn"] = None
    _registry.store(path, res)
    th = np.arange(0.0, 181.0, 2.0)
    nf.CalcNF2FF(path, F.exc[1] if args.workload != "patch100m" else 2.45e9, th, np.array([0.0]), center=[0, 0, 0])
    if getattr(sim, "td_store", False):
        nf.CalcNF2FF(path, 0.97 * F.exc[1], th, np.array([90.0]), center=[0, 0, 0])     # unregistered: DFT of the stored samples
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"profiled {args.steps + 1} steps of {args.workload}: grid {sim.nx}x{sim.ny}x{sim.nz}, interval {iv}, he_active {E.he_active}, "
      f"plan {E.plan_info()}, energy {e:.3e}")
