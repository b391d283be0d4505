"""GPU-side sweep of the volume-kernel tuning knobs (kz, ty) on a synthetic cube.  Dev tool."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fdtd-solver-antennas_b200")]
import torch  # noqa: E402
from b200fdtd.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--nz", type=int, default=0)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--kz", type=str, default="4,8,16,32,64")
ap.add_argument("--ty", type=str, default="2,4,8")
ap.add_argument("--variant", type=str, default="0")
args = ap.parse_args()
n = args.n
nz = args.nz or n
E = Engine(n, n, nz, px=(n + 31) // 32 * 32)
g = torch.Generator(device="cuda").manual_seed(0)
shape = E.shape
co = [torch.rand(shape, device="cuda", generator=g) * 0.1 + 0.9 for _ in range(2)] + \
     [torch.rand(shape, device="cuda", generator=g) * 0.1 + 0.05 for _ in range(2)]
E.set_coeffs(co[0], co[2], co[1], co[3])
E.volt.normal_(generator=g)
E.curr.normal_(generator=g)
cells = n * n * nz
res = []
for variant in [int(v) for v in args.variant.split(",")]:
    for ty in [int(v) for v in args.ty.split(",")]:
        for kz in [int(v) for v in args.kz.split(",")]:
            E.set_tuning(kz=kz, ty=ty, variant=variant)
            for which in (0, 1):
                for _ in range(3):
                    E.update_only(which, join=False)
                E._pre(); torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(E.stream)
                for _ in range(args.iters):
                    E.update_only(which, join=False)
                b.record(E.stream)
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / args.iters
                gbs = cells * 60 / ms / 1e6
                res.append(dict(variant=variant, ty=ty, kz=kz, which="EH"[which], ms=round(ms, 4), GBs=round(gbs, 1)))
                print(res[-1], flush=True)
best = {}
for r in res:
    k = r["which"]
    if k not in best or r["GBs"] > best[k]["GBs"]:
        best[k] = r
print("BEST", json.dumps(best))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"sweep_{n}x{nz}.json"), "w"), indent=1)
