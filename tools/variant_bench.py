"""GPU-side A/B of engine variants (b200fdtd_set_tuning) on one scene: whole-step time by CUDA graph replay.  Dev tool.

  python tools/variant_bench.py --variants 0,32 --steps 100 [--workload cube --cube-n 768]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fdtd-solver-antennas_b200")]
import torch  # noqa: E402
from b200fdtd import scenes  # noqa: E402
from b200fdtd.simulation import Simulation  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--variants", default="0")
ap.add_argument("--kz", default="16")
ap.add_argument("--ty", default="4")
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--cells", type=float, default=100e6)
ap.add_argument("--workload", default="patch100m")
ap.add_argument("--cube-n", type=int, default=768)
ap.add_argument("--out", default="variants.json")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--he", default="0:0", help="fused H->E tiles rows:planes[:lookahead][,rows:planes...] (0 = keep / automatic)")
args = ap.parse_args()

if args.workload == "cube":
    F = scenes.vacuum_cube(args.cube_n, nrts=10 ** 6)
else:
    F, nf, port = scenes.patch_scene(target_cells=args.cells, boundary="PML_8", f0=2.5e9, fc=1.5e9, nrts=10 ** 6,
                                     end_criteria=1e-12, nf2ff_freqs=[2.45e9])
S = F._setup()
sim = Simulation(S, device=0, nf2ff_freqs=F.nf2ff_freqs, probe_freqs=S.probe_freqs)
sim.prepare()
E = sim.engine
res = []
for he in args.he.split(","):
  hv = [int(v) for v in he.split(":")] + [0]
  E.set_he_tuning(hv[0], hv[1], de=hv[2])
  for variant in [int(v) for v in args.variants.split(",")]:
    for ty in [int(v) for v in args.ty.split(",")]:
        for kz in [int(v) for v in args.kz.split(",")]:
            E.set_tuning(kz=kz, ty=ty, variant=variant)
            E.run(sim.interval * 2, use_graph=True)
            torch.cuda.synchronize()
            best = 1e9
            for rep in range(args.reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                E._pre()
                a.record(E.stream)
                E.run(args.steps, use_graph=True)
                b.record(E.stream)
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b) / args.steps)
            ms = best
            kms = None
            if E.he_active:                      # the fused launch alone
                E.update_only(4, join=False)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); a.record(E.stream)
                for _ in range(10):
                    E.update_only(4, join=False)
                b.record(E.stream); torch.cuda.synchronize()
                kms = round(a.elapsed_time(b) / 10, 4)
            r = dict(he=he, fused=E.he_active, variant=variant, ty=ty, kz=kz, ms_per_step=round(ms, 5), mcells=round(sim.cells / ms / 1e3, 1), he_kernel_ms=kms)
            res.append(r)
            print(r, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dict(grid=[sim.nx, sim.ny, sim.nz_glob], results=res), open(os.path.join(ROOT, "gpurun_out", args.out), "w"), indent=1)
