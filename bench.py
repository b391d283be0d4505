#!/usr/bin/env python
"""bench.py — FDTD Mcell-updates/s of the B200 engine on the BASELINE.json workload.

  python bench.py --gpus N --steps K --warmup W            (own arm; torchrun launches N ranks for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (the CPU engine on the box's host cores)

Workload (config.workload):
  N = 1 : "patch100m" = BASELINE.json configs[1]: the 2.45 GHz FR-4 patch scene on a refined ~100 M-cell mesh with
          PML_8 and a 1-4 GHz Gaussian pulse (f0 2.5 GHz, fc 1.5 GHz), lumped port, port V/I probes and DFT, NF2FF box DFT.
  N > 1 : the same scene with the mesh refined so every rank keeps ~100 M cells (weak scaling), z-slab sharded.
One "step" = one full FDTD time step (E and H update of every cell plus all boundary / port / DFT kernels).
Prints ONE JSON line (see DESIGN.md §Measurement for every field).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "fdtd-solver-antennas_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# stdout carries exactly one JSON line: NCCL's own log lines (e.g. the version banner some images switch on through
# NCCL_DEBUG) go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
    os.environ.pop("NCCL_DEBUG")

import numpy as np  # noqa: E402

BYTES_PER_CELL_PASS = 60          # 15 fp32 streams per E (or H) pass
BYTES_PER_CELL_STEP = 120


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="patch100m", choices=["patch100m", "cube"])
    ap.add_argument("--cells", type=float, default=100e6, help="target cells PER GPU for the patch workload")
    ap.add_argument("--cube-n", dest="n", type=int, default=512, help="cube edge (workload cube)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kz", type=int, default=0, help="tuning experiment: planes marched per CTA")
    ap.add_argument("--ty", type=int, default=0, help="tuning experiment: rows per CTA")
    ap.add_argument("--variant", type=int, default=0, help="tuning experiment: engine variant bits")
    ap.add_argument("--no-align-x", action="store_true", help="tuning experiment: minimal (unaligned) x-slab PML boxes")
    ap.add_argument("--cpu-cells", type=float, default=12.5e6, help="cells of the CPU-baseline sample of the same scene")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------- helpers
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def build_scene(args, world, target_cells):
    from b200fdtd import scenes
    if args.workload == "cube":
        return scenes.vacuum_cube(args.n, nrts=10 ** 6), None, None
    F, nf, port = scenes.patch_scene(target_cells=target_cells * world, boundary="PML_8", f0=2.5e9, fc=1.5e9,
                                     nrts=10 ** 6, end_criteria=1e-12, nf2ff_freqs=[2.45e9])
    return F, nf, port


# ---------------------------------------------------------------------------------------------- CPU engine (oracle)
def cpu_engine_run(args, steps, warmup, cells):
    """time the CPU engine (oracle/fdtd_ref.c, OpenMP over all host cores) on the same scene recipe at `cells` cells"""
    import torch
    from oracle.fdtd_ref import RefEngine, lib
    from openEMS import openEMS as O
    threads = os.cpu_count() or 1
    saved = (O.default_engine_factory, O.default_farfield_fn)
    O.default_engine_factory = staticmethod(lambda nx, ny, nz, px, dev: RefEngine(nx, ny, nz, px, threads=threads))
    try:
        F, nf, port = build_scene(args, 1, cells)
        S = F._setup()
        from b200fdtd.simulation import Simulation
        sim = Simulation(S, device=0, engine_factory=F.engine_factory, build_device=torch.device("cpu"), px_align=8,
                         nf2ff_freqs=F.nf2ff_freqs, probe_freqs=S.probe_freqs)
        sim.prepare()
        E = sim.engine
        E.run(warmup)
        t0 = time.perf_counter()
        E.run(steps)
        dt = time.perf_counter() - t0
    finally:
        O.default_engine_factory, O.default_farfield_fn = saved
    ncell = sim.cells
    return dict(value=ncell * steps / dt / 1e6, seconds=dt, cells=ncell, threads=threads, steps=steps,
                grid=[sim.nx, sim.ny, sim.nz_glob], lib_threads=int(lib().ref_max_threads()))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    r = cpu_engine_run(args, steps, max(1, warmup), args.cpu_cells)
    sample = (f"{steps} time steps of the same patch scene recipe meshed at {r['cells']} cells "
              f"({r['grid'][0]}x{r['grid'][1]}x{r['grid'][2]}, PML_8, port, probes, NF2FF DFT)")
    line = {
        "impl": "reference", "metric": "FDTD Mcell-updates/s", "value": round(r["value"], 2), "unit": "Mcell/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": round(1e3 * r["seconds"] / steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "patch100m" if args.workload == "patch100m" else f"cube{args.n}",
                   "engine": "openEMS-convention CPU engine (restated, oracle/fdtd_ref.c; real openEMS is not installable here)",
                   "sample_cells": r["cells"]},
        "cpu_baseline": {"value": round(r["value"], 2), "unit": "Mcell/s", "cores": r["threads"], "kind": "port", "sample": sample},
        "e2e": {"value": round(r["value"], 2), "unit": "Mcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- own arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from b200fdtd import engine as eng_mod
    from b200fdtd.simulation import Simulation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, W = args.steps, max(3, args.warmup)

    F, nf, port = build_scene(args, world, args.cells)
    S = F._setup()
    t0 = time.time()
    sim = Simulation(S, device=local, rank=rank, world=world, nf2ff_freqs=F.nf2ff_freqs, probe_freqs=S.probe_freqs,
                     align_x_slabs=not args.no_align_x)
    sim.prepare()
    build_s = time.time() - t0
    E = sim.engine
    if args.kz or args.ty or args.variant:
        E.set_tuning(kz=args.kz or 16, ty=args.ty or 4, variant=args.variant)
    cells = sim.cells
    local_cells = sim.nx * sim.ny * sim.nz

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(n):
        if world > 1:
            sim._step_multi(n)
        else:
            E.run(n, use_graph=True)

    # ---- resident throughput: K full steps, inputs already in HBM ----
    step(W)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = eng_mod.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    E._pre()
    ev0.record(E.stream)
    step(K)
    ev1.record(E.stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng_mod.launch_count() - l0
    # ---- dominant kernel, timed alone on the same data, same stream ----
    # the fused H->E launch (H update of step n + E update of step n+1 of the plain region in one sweep, 120 B/cell
    # algorithmic); if the run could not use it (no room for the second field copy): the separate plain E and H launches
    reps = 10
    kms = []
    plain_cells, fused_cells, sep_cells = E.plan_info()
    fused_he = (world == 1 and E.he_active) or (world > 1 and bool(getattr(sim, "_fused", False)))
    for which in ((4,) if fused_he else (2, 3)):
        E.update_only(which, join=False)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(E.stream)
        for _ in range(reps):
            E.update_only(which, join=False)
        b.record(E.stream)
        torch.cuda.synchronize()
        kms.append(a.elapsed_time(b) / reps)
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        t = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        launches = int(t.item())
    value = cells * K / (ms / 1e3) / 1e6

    # ---- end to end with HOST buffers: operator H2D + K steps + results D2H ----
    # The host holds the operator in the form the host builder produces it (x-vector tables + 32-byte row records + the
    # few non-separable rows, pinned); the timed region uploads it, expands it into the bound arrays on the device,
    # re-verifies the compression (C-ABI b200fdtd_set_row_compression), steps K times and reads every result back.
    host_op = sim.export_operator(pin=True)
    big = local_cells > 300e6              # no room for a second copy of the operator: compare fp64 checksums + sample planes
    if big:
        before = [(float(torch.sum(t, dtype=torch.float64)), t[:, 1::max(1, sim.nz // 7)].clone()) for t in (E.vv, E.vi, E.ii, E.iv)]
    else:
        before = [t.clone() for t in (E.vv, E.vi, E.ii, E.iv)]
    for t in (E.vv, E.vi, E.ii, E.iv):
        t.zero_()
    torch.cuda.synchronize()
    outs = [o for o in [E.series, E.probe_dft] + list(E.face_acc) if o is not None]
    res_host = [torch.empty(o.shape, dtype=o.dtype, pin_memory=True) for o in outs]    # pinned result buffers on the host
    barrier()
    t_e0 = time.perf_counter()
    sim.load_operator(host_op)
    step(K)
    for o, h in zip(outs, res_host):
        h.copy_(o, non_blocking=True)
    barrier()
    e2e_s = time.perf_counter() - t_e0
    if big:
        e2e_ok = [int(float(torch.sum(t, dtype=torch.float64)) != cs) + int((t[:, 1::max(1, sim.nz // 7)] != smp).count_nonzero())
                  for (cs, smp), t in zip(before, (E.vv, E.vi, E.ii, E.iv))]
    else:
        e2e_ok = [sum(int((a[c] != b[c]).count_nonzero()) for c in range(3)) for a, b in zip(before, (E.vv, E.vi, E.ii, E.iv))]
    del before
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = sim.operator_nbytes(host_op)
    d2h = sum(o.numel() * o.element_size() for o in res_host)
    e2e_value = cells * K / e2e_s / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    if fused_he:
        k_ms, bytes_per_cell_launch = kms[0], BYTES_PER_CELL_STEP
        kname = "update_he5_kernel<7> fused H->E launch over the plain region (one launch = both passes)"
        kernel_ms = {"HE": round(kms[0], 4)}
    else:
        k_ms, bytes_per_cell_launch = 0.5 * (kms[0] + kms[1]), BYTES_PER_CELL_PASS
        kname = "update_e_kernel<4,0,1>/update_h_kernel<4,0,1> plain launch (mean of both passes)"
        kernel_ms = {"E": round(kms[0], 4), "H": round(kms[1], 4)}
    achieved = bytes_per_cell_launch * plain_cells / (k_ms / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            per_cell = tj.get("dram_bytes_per_cell_fused_launch" if fused_he else "dram_bytes_per_cell_pass")
            traffic = None if per_cell is None else per_cell * plain_cells
        except Exception:
            traffic = None
    line = {
        "metric": "FDTD Mcell-updates/s", "value": round(value, 1), "unit": "Mcell/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms / K, 5), "higher_is_better": True, "scaling": "weak" if args.workload == "patch100m" else "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "patch100m" if args.workload == "patch100m" else f"cube{args.n}",
                   "scene": ("2.45 GHz FR-4 patch (reference recipe), PML_8, Gaussian 1-4 GHz, lumped port, V/I probes + DFT, NF2FF DFT"
                             if args.workload == "patch100m" else "uniform vacuum cube, Mur on 6 faces, centre soft source, one V probe (config 5)"),
                   "grid": [sim.nx, sim.ny, sim.nz_glob], "cells": cells, "cells_per_gpu": local_cells, "pml_cells_rank0": sim.pml_cells,
                   "timestep_s": sim.dt, "sample_interval": sim.interval, "parallelism": f"z-slab x{world}",
                   "l2_note": "working set 72 B/cell >> 126 MB L2 (inputs larger than L2, no flush needed)",
                   "operator_build_s": round(build_s, 2),
                   "row_compression": {("E" if w == 0 else "H"): {"row_slots_compressed": v[0], "row_slots_demoted": v[1], "x_vectors": v[2]}
                                       for w, v in sim.compression.items()}},
        "roofline": {"bound": "hbm", "kernel": kname,
                     "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": traffic, "peak_source": peak_src, "kernel_ms": kernel_ms,
                     "bytes_per_cell_launch": bytes_per_cell_launch, "cells_per_launch": plain_cells,
                     "fused_pml_cells": fused_cells, "separate_pml_cells": sep_cells,
                     "whole_step_frac": round(BYTES_PER_CELL_STEP * local_cells / (ms / K / 1e3) / 1e9 / peak, 4),
                     "frac_of_nominal_8TBs": round(achieved / 8000.0, 4)},
        "e2e": {"value": round(e2e_value, 1), "unit": "Mcell/s", "h2d_bytes_per_step": int(h2d / K), "d2h_bytes_per_step": int(d2h / K),
                "what": "compressed operator (x-vector tables, row records, non-separable rows) from pinned host memory -> device, "
                        "expanded + verified on the device, K steps, probe series/DFT + NF2FF spectra -> host",
                "operator_restored_mismatches": e2e_ok,
                "seconds": round(e2e_s, 4)},
        "gpu_launches": int(launches),
        "clocks": clk,
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            # bounded sample: ~10-30 s of CPU work on the same scene recipe at a reduced mesh
            r = cpu_engine_run(args, 12, 2, args.cpu_cells)
            line["cpu_baseline"] = {"value": round(r["value"], 2), "unit": "Mcell/s", "cores": r["threads"], "kind": "port",
                                    "sample": f"12 time steps of the same scene recipe meshed at {r['cells']} cells "
                                              f"({r['grid'][0]}x{r['grid'][1]}x{r['grid'][2]}); oracle/fdtd_ref.c with OpenMP"}
        except Exception as e:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": "Mcell/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
