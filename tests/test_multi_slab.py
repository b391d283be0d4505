"""z-slab decomposition: 2 ranks over gloo on the CPU reproduce the single-slab run.

Uses the oracle engine as the per-rank engine (CPU box has no GPU); the host logic under test — slab ranges,
per-slab operator build, index remapping, halo exchange order, result reduction — is the product's.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _scene(scene, boundary):
    import scenes
    if scene == "dipole":
        F, nf, prt = scenes.dipole(boundary, cells=(20, 20, 30), nrts=300, end=1e-12)
        return F
    # BASELINE.json configs[2]: the 4x4 array emitted by the unmodified multi prepare (recorded trace), cut to a few steps
    import replay
    F = replay.replay(scene)["FDTD"]
    F.SetNumberOfTimeSteps(int(os.environ.get("B200FDTD_TEST_ARRAY_STEPS", "190")))
    F.SetEndCriteria(1e-30)
    return F


def _worker(rank, world, port, boundary, fused, q, scene="dipole"):
    sys.path[:0] = [os.path.dirname(HERE), HERE, os.path.join(os.path.dirname(HERE), "fdtd-solver-antennas_b200")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scenes
        scenes.use_oracle_engine(threads=1 if scene == "dipole" else 4)
        F = _scene(scene, boundary)
        F.fused_multi = fused
        F.Run(scenes.tmp_sim_path(f"slab{rank}"), cleanup=True)
        assert F.sim._fused == fused, "the run did not take the requested stepping protocol"
        res = F.results
        if rank == 0:
            q.put(dict(probes={k: v["val"] for k, v in res["probes"].items()},
                       acc=[a.copy() for a in res["nf2ff"]["acc"]], dt=res["dt"], ts=res["timesteps"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("boundary,world,fused,scene", [
    ("PML_8", 2, True, "dipole"), ("PML_8", 3, True, "dipole"), ("PML_8", 2, False, "dipole"),
    ("MUR", 2, True, "dipole"), ("MUR", 3, True, "dipole"), ("MUR", 2, False, "dipole"),
    ("MUR", 2, True, "trace_array16_mur_q1")])
def test_slabs_equal_single(boundary, world, fused, scene):
    """fused=True: the fused-step protocol of the CUDA engine (simulation.py:_fused_step; the oracle engine emulates the two
    field copies and poisons the stale one with NaN); fused=False: separate half steps with one exchange each.
    The last case is BASELINE.json configs[2]: the 16-element array (16 ports, rotated/translated boxes) cut in two slabs."""
    import scenes
    scenes.use_oracle_engine(threads=1 if scene == "dipole" else 8)
    F = _scene(scene, boundary)
    F.Run(scenes.tmp_sim_path("slab_single"), cleanup=True, distributed=False)
    ref = F.results
    scenes.use_cuda_engine()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world + (10 if fused else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, boundary, fused, q, scene)) for r in range(world)]
    for p in procs:
        p.start()
    got = None
    for _ in range(600):
        if not q.empty():
            got = q.get()
            break
        if any(p.exitcode not in (None, 0) for p in procs):
            break
        procs[0].join(timeout=0.5)
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
    assert got is not None and all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert got["dt"] == ref["dt"] and got["ts"] == ref["timesteps"]
    assert sorted(got["probes"]) == sorted(ref["probes"]) and len(got["probes"]) >= 2
    for name, a in got["probes"].items():
        b = ref["probes"][name]["val"]
        assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max(), name
    for a, b in zip(got["acc"], ref["nf2ff"]["acc"]):
        assert np.abs(a - b).max() <= 1e-12 * max(np.abs(b).max(), 1e-300)
