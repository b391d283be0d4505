"""torchrun entry: z-slab run on N GPUs (NCCL halo exchange, overlapped) must equal the 1-GPU run.  Prints PASS/FAIL."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "fdtd-solver-antennas_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import scenes  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for boundary in ("PML_8", "MUR"):
        F, nf, port = scenes.dipole(boundary, cells=(56, 48, 64), nrts=600, end=1e-12)
        F.device = local
        F.Run(scenes.tmp_sim_path(f"mg{rank}"), cleanup=True)
        multi = F.results
        if rank == 0:
            F1, nf1, port1 = scenes.dipole(boundary, cells=(56, 48, 64), nrts=600, end=1e-12)
            F1.device = local
            F1.Run(scenes.tmp_sim_path("mg_single"), cleanup=True, distributed=False)
            single = F1.results
            for name in ("port_ut_1", "port_it_1"):
                a, b = multi["probes"][name]["val"], single["probes"][name]["val"]
                err = np.abs(a - b).max() / np.abs(b).max()
                print(f"{boundary} {name}: rel err {err:.3e} over {len(b)} samples")
                ok &= bool(err < 1e-5)
            for fa, fb in zip(multi["nf2ff"]["acc"], single["nf2ff"]["acc"]):
                err = np.abs(fa - fb).max() / np.abs(fb).max()
                ok &= bool(err < 1e-4)
            print(f"{boundary} nf2ff faces max rel err ok={ok}; dt equal: {multi['dt'] == single['dt']}")
            ok &= multi["dt"] == single["dt"]
        dist.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", f"world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
