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
        macc = multi["nf2ff"]["acc"]                 # host copy of the face spectra: a collective on z-slab runs, every rank asks
        th, ph = np.arange(0.0, 181.0, 15.0), np.array([0.0, 90.0])
        ff_multi = nf.CalcNF2FF(scenes.tmp_sim_path(f"mg{rank}"), F.exc[1], th, ph)      # far-field sums added over the ranks
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
            for fa, fb in zip(macc, single["nf2ff"]["acc"]):
                err = np.abs(fa - fb).max() / np.abs(fb).max()
                ok &= bool(err < 1e-4)
            ff_single = nf1.CalcNF2FF(scenes.tmp_sim_path("mg_single"), F1.exc[1], th, ph)
            e_err = np.abs(ff_multi.E_norm[0] - ff_single.E_norm[0]).max() / ff_single.E_norm[0].max()
            d_err = abs(ff_multi.Dmax[0] / ff_single.Dmax[0] - 1.0)
            print(f"{boundary} far field from per-rank sources: E_norm rel err {e_err:.2e}, Dmax rel err {d_err:.2e}")
            ok &= bool(e_err < 1e-4 and d_err < 1e-4)
            print(f"{boundary} nf2ff faces max rel err ok={ok}; dt equal: {multi['dt'] == single['dt']}")
            ok &= multi["dt"] == single["dt"]
        dist.barrier()
    # BASELINE.json configs[2]: the 4x4 array of the unmodified multi prepare (recorded trace, 16 ports), cut to 400 steps
    import replay
    def array16():
        F = replay.replay("trace_array16_mur_q1")["FDTD"]
        F.SetNumberOfTimeSteps(400); F.SetEndCriteria(1e-30); F.device = local
        return F
    F = array16()
    F.Run(scenes.tmp_sim_path(f"mg16_{rank}"), cleanup=True)
    multi = F.results
    macc = multi["nf2ff"]["acc"]
    if rank == 0:
        F1 = array16()
        F1.Run(scenes.tmp_sim_path("mg16_single"), cleanup=True, distributed=False)
        single = F1.results
        worst = 0.0
        for name in sorted(single["probes"]):
            a, b = multi["probes"][name]["val"], single["probes"][name]["val"]
            worst = max(worst, float(np.abs(a - b).max() / np.abs(b).max()))
        for fa, fb in zip(macc, single["nf2ff"]["acc"]):
            worst = max(worst, float(np.abs(fa - fb).max() / np.abs(fb).max()))
        print(f"array16 (16 ports, {len(single['probes'])} probes): worst rel err {worst:.3e}; fused multi-GPU steps: {F.sim._fused}; dt equal: {multi['dt'] == single['dt']}")
        ok &= bool(worst < 1e-4) and multi["dt"] == single["dt"]
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", f"world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
