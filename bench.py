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
    ap.add_argument("--steps", type=int, default=248)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="patch100m", choices=["patch100m", "cube"])
    ap.add_argument("--cells", type=float, default=100e6, help="target cells PER GPU for the patch workload")
    ap.add_argument("--cube-n", dest="n", type=int, default=512, help="cube edge (workload cube)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-seeded", action="store_true", help="skip the seeded per-cell-coefficient record (N = 1)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling record (1024^3 Mur cube)")
    ap.add_argument("--no-latency", action="store_true", help="skip the small-scene end-to-end record (N = 1)")
    ap.add_argument("--no-config3", action="store_true", help="skip the 4x4 array record (N > 1)")
    ap.add_argument("--strong-n", type=int, default=1024, help="edge of the strong-scaling cube")
    ap.add_argument("--strong-nz", type=int, default=0, help="experiment: z extent of the strong-scaling grid (0 = cube)")
    ap.add_argument("--config3-cells", type=float, default=1.0e9, help="total cells of the 4x4 array mesh")
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


def build_scene(args, world, target_cells, nrts=10 ** 6):
    from b200fdtd import scenes
    if args.workload == "cube":
        return scenes.vacuum_cube(args.n, nrts=nrts), None, None
    F, nf, port = scenes.patch_scene(target_cells=target_cells * world, boundary="PML_8", f0=2.5e9, fc=1.5e9,
                                     nrts=nrts, end_criteria=1e-12, nf2ff_freqs=[2.45e9])
    return F, nf, port


def build_array16(target_cells, nrts):
    """BASELINE.json configs[2]: the 4x4 patch array emitted by the UNMODIFIED reference multi prepare
    (antenna_sim/solver_fdtd_openems_microstrip_multi_3d.py:98-593; recorded call trace tests/golden/trace_array16_pml8_q2.json,
    16 PatchInstances at 60 mm pitch, PML_8, 16 lumped ports fed in phase), then refined through the public grid API
    (SmoothMeshLines with a smaller max_res) until the mesh holds ~target_cells cells"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import replay
    from b200fdtd import mesh as _mesh
    R = replay.replay("trace_array16_pml8_q2")
    F = R["FDTD"]
    grid = F.GetCSX().GetGrid()
    base = [grid.GetLines(a, do_sort=True) for a in range(3)]

    def cells(res):
        return float(np.prod([len(_mesh.smooth_mesh_lines(l, res, 1.4)) for l in base]))
    lo, hi = 0.02, max(float(np.diff(l).max()) for l in base)
    for _ in range(40):
        mid = (lo * hi) ** 0.5
        if cells(mid) > target_cells:
            lo = mid
        else:
            hi = mid
        if hi / lo < 1.002:
            break
    grid.SmoothMeshLines("all", hi, 1.4)
    F.SetNumberOfTimeSteps(nrts)
    F.SetEndCriteria(1e-30)
    F.nf2ff_td = False
    return F, R["nf"], F.ports[0], hi


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
class Runner:
    """one prepared scene on this rank: resident stepping, kernel timing"""

    def __init__(self, F, local, world, args):
        import torch
        self.torch = torch
        self.F, self.world = F, world
        F.device = local
        t0 = time.time()
        # the public entry: FDTD.Run(sim_path, setup_only=True) builds the operator and keeps it on the host in the form the
        # host builder produces it; a later FDTD.Run of the same scene ships it to the device
        self.path = os.path.join("/tmp", f"b200fdtd_bench_{os.getpid()}")
        F.Run(self.path, setup_only=True, cleanup=True)
        self.build_s = time.time() - t0
        self.sim = F._prepared[1]
        self.E = self.sim.engine
        if args.kz or args.ty or args.variant:
            self.E.set_tuning(kz=args.kz or 16, ty=args.ty or 4, variant=args.variant)

    def barrier(self):
        if self.world > 1:
            self.torch.distributed.barrier()
        self.torch.cuda.synchronize()

    def step(self, n):
        if n <= 0:
            return
        if self.world > 1:
            self.sim._step_multi(n)
        else:
            self.E.run(n, use_graph=True)

    def resident(self, K, W):
        """K timed steps after >= W warm-up steps.  The start is aligned so that the timed region holds what a long run holds:
        K >= sampling interval: starts on an interval boundary (graph replay of whole intervals, each ending in the probe /
        NF2FF sampling launches); K < interval: the one sampling point falls in the middle of the K steps."""
        torch, E, sim = self.torch, self.E, self.sim
        from b200fdtd import engine as eng_mod
        iv = sim.interval
        self.step(W)
        target = 0 if K >= iv else (iv - (K + 1) // 2) % iv
        self.step((target - E.ts) % iv)
        if K >= iv and self.world == 1:
            # the same call once untimed: CUDA graphs are captured lazily per state of the field copies at their entry, and a
            # capture (hundreds of kernel nodes, tens of milliseconds on the host) must not fall into the timed region
            self.step(K)
            self.step((target - E.ts) % iv)
        ts0 = E.ts
        self.barrier()
        l0 = eng_mod.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        E._pre()
        ev0.record(E.stream)
        t_h = time.perf_counter()
        self.step(K)
        host_ms = 1e3 * (time.perf_counter() - t_h)      # time the host needed to enqueue the K steps (not a GPU time)
        ev1.record(E.stream)
        self.barrier()
        ms = ev0.elapsed_time(ev1)
        launches = eng_mod.launch_count() - l0
        samplings = (ts0 + K) // iv - ts0 // iv
        # one fused span per call: K - 1 fused steps, replayed from graphs of g steps (g = interval, twice that if odd)
        g = iv if iv % 2 == 0 else 2 * iv
        graph_steps = ((K - 1) // g) * g if (self.world == 1 and K >= iv and getattr(E, "he_active", False)) else 0
        if self.world > 1:
            t = torch.tensor([ms, float(launches)], dtype=torch.float64, device="cuda")
            m = t.clone()
            torch.distributed.all_reduce(m, op=torch.distributed.ReduceOp.MAX)
            torch.distributed.all_reduce(t)
            ms, launches = float(m[0].item()), int(t[1].item())
        return ms, int(launches), dict(start_ts=int(ts0), sampling_launch_sets=int(samplings), graph_replayed_steps=int(graph_steps),
                                       eager_steps=int(K - graph_steps), host_enqueue_ms_per_step=round(host_ms / K, 4))

    def kernel_times(self, reps=10):
        """the dominant kernel alone, same data, same stream: the fused H->E launch if the run used it, else the plain E and H launches"""
        torch, E = self.torch, self.E
        fused = (self.world == 1 and E.he_active) or (self.world > 1 and bool(getattr(self.sim, "_fused", False)))
        kms = []
        for which in ((4,) if fused else (2, 3)):
            E.update_only(which, join=False)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(E.stream)
            for _ in range(reps):
                E.update_only(which, join=False)
            b.record(E.stream)
            torch.cuda.synchronize()
            kms.append(a.elapsed_time(b) / reps)
        return fused, kms


def traffic_per_cell(key):
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(tp)).get(key)
    except Exception:
        return None


def roofline_record(fused, kms, plain_cells, peak, peak_src, kname_fused, kname_split, traffic_keys):
    if fused:
        k_ms, bpc, kname, kernel_ms = kms[0], BYTES_PER_CELL_STEP, kname_fused, {"HE": round(kms[0], 4)}
        per_cell = traffic_per_cell(traffic_keys[0])
    else:
        k_ms, bpc, kname = 0.5 * (kms[0] + kms[1]), BYTES_PER_CELL_PASS, kname_split
        kernel_ms = {"E": round(kms[0], 4), "H": round(kms[1], 4)}
        per_cell = traffic_per_cell(traffic_keys[1])
    achieved = bpc * plain_cells / (k_ms / 1e3) / 1e9
    traffic = None if per_cell is None else per_cell * plain_cells
    rec = {"bound": "hbm", "kernel": kname, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
           "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src, "kernel_ms": kernel_ms,
           "bytes_per_cell_launch": bpc, "cells_per_launch": plain_cells, "frac_of_nominal_8TBs": round(achieved / 8000.0, 4)}
    if traffic is not None:
        # what the kernel really moves (ncu dram__bytes per launch, profiles/traffic.json) over its live CUDA-event time
        rec["actual_dram_gbs"] = round(traffic / (k_ms / 1e3) / 1e9, 1)
        rec["actual_frac"] = round(traffic / (k_ms / 1e3) / 1e9 / peak, 4)
        rec["traffic_source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch (profiles/traffic.json), scaled by cells"
    return rec


def seeded_record(R, K, W, peak, peak_src):
    """SURVEY.md §8(d) guard run: the same grid with a SEEDED per-cell operator (torch.manual_seed(0): vv, ii scaled by
    U[0.9,1], vi, iv by U[0.5,1] cell by cell), so no row compresses and nothing can be constant-folded: the per-cell
    coefficient kernels the north star describes, 120 B/cell of real traffic"""
    torch, E, sim = R.torch, R.E, R.sim
    g = torch.Generator(device=E.device).manual_seed(0)
    for arr, lo in ((E.vv, 0.9), (E.ii, 0.9), (E.vi, 0.5), (E.iv, 0.5)):
        for c in range(3):
            for k0 in range(0, arr.shape[1], 32):
                blk = arr[c, k0:k0 + 32]
                blk.mul_(torch.empty_like(blk).uniform_(lo, 1.0, generator=g))
    for which in (0, 1):
        E.set_row_compression(which, None, None)
    E.reset_state()
    ms, launches, region = R.resident(K, W)
    fused, kms = R.kernel_times()
    plain_cells = E.plan_info()[0]
    rec = {"what": "same grid, seeded per-cell coefficients (torch.manual_seed(0)), row compression off",
           "value": round(sim.cells * K / (ms / 1e3) / 1e6, 1), "unit": "Mcell/s", "ms_per_step": round(ms / K, 5),
           "gpu_launches": launches, "timed_region": region,
           "roofline": roofline_record(fused, kms, plain_cells, peak, peak_src,
                                       "fused H->E launch, per-cell coefficient streams (one launch = both passes)",
                                       "update_e_kernel<4,0,0>/update_h_kernel<4,0,0> plain launch, per-cell coefficients (mean of both passes)",
                                       ("dram_bytes_per_cell_fused_launch_uncompressed", "dram_bytes_per_cell_pass_uncompressed"))}
    return rec


def latency_record(args, local):
    """BASELINE.json configs[0] end to end through the public API: the reference's own single-patch scene (recorded call trace
    of the unmodified prepare, GUI defaults: PML_8, mesh quality 3, NrTS 30000, EndCriteria 1e-4, ~0.3 M cells), FDTD.Run to
    its end criterion + CalcPort + the reference's 73-call CalcNF2FF loop; timed on the second run of the process (the first
    one pays CUDA context creation and module loading), next to the CPU engine on all host cores"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import replay
    import scenes as tscenes
    out = {"workload": "patch_q3", "scene": "reference single-patch scene, unmodified prepare (trace_single_pml8_q3): PML_8, NrTS 30000, EndCriteria 1e-4"}

    def once(tag):
        R = replay.replay("trace_single_pml8_q3")
        F, nf = R["FDTD"], R["nf"]
        F.device = local
        path = os.path.join("/tmp", f"b200fdtd_lat_{tag}_{os.getpid()}")
        t0 = time.perf_counter()
        F.Run(path, cleanup=True, verbose=0)
        t_run = time.perf_counter() - t0
        f = np.linspace(max(1e9, 0.7 * F.exc[1]), 1.3 * F.exc[1], 201)
        F.ports[0].CalcPort(path, f)
        dbi, dmax = replay.reference_postprocess(nf, path, F.exc[1], R["theta"], R["phi"], R["nf_center"])
        t_all = time.perf_counter() - t0
        sim = F.sim
        return dict(run_s=round(t_run, 4), total_s=round(t_all, 4), prepare_s=round(sim.prepare_s, 4), step_loop_s=round(sim.wall_s, 4),
                    timesteps=sim.timesteps, stop=sim.stop_reason, cells=sim.cells, us_per_step=round(1e6 * sim.wall_s / sim.timesteps, 2),
                    mcells_per_s=round(sim.cells * sim.timesteps / sim.wall_s / 1e6, 1), dmax_dbi=round(float(10 * np.log10(dmax)), 4))
    tscenes.use_cuda_engine()
    out["cold"] = once("cold")
    out["warm"] = once("warm")
    if not args.no_cpu_baseline:
        tscenes.use_oracle_engine(threads=os.cpu_count() or 4)
        try:
            out["cpu"] = once("cpu")
            out["cpu"]["cores"] = os.cpu_count()
            out["speedup_total_warm"] = round(out["cpu"]["total_s"] / out["warm"]["total_s"], 2)
            out["speedup_run_warm"] = round(out["cpu"]["run_s"] / out["warm"]["run_s"], 2)
        finally:
            tscenes.use_cuda_engine()
    return out


def sub_workload(kind, args, local, world, K, W, peak, peak_src):
    """further records of the same line: 'strong' = BASELINE configs[4] 1024^3 vacuum cube with Mur, total size fixed as N grows;
    'config3' = BASELINE configs[2] 4x4 array at ~1 B cells over N GPUs"""
    import copy
    import torch
    a = copy.copy(args)
    t0 = time.time()
    if kind == "strong":
        from b200fdtd import scenes
        F = scenes.vacuum_cube(args.strong_n, nrts=10 ** 6, nz=args.strong_nz or None)
        desc = f"uniform vacuum cube {args.strong_n}^3, Mur on 6 faces, centre soft source, one V probe (config 5); total size fixed (strong scaling)"
        extra = {}
    else:
        F, nf, port, res = build_array16(args.config3_cells, 10 ** 6)
        desc = ("4x4 patch array from the unmodified multi prepare (recorded trace), mesh refined with SmoothMeshLines to "
                f"max_res {res:.4f} mm, PML_8, 16 lumped ports, probes + DFT, NF2FF DFT; total size fixed")
        extra = {"max_res_mm": round(res, 5)}
    R = Runner(F, local, world, a)
    ms, launches, region = R.resident(K, W)
    fused, kms = R.kernel_times()
    sim, E = R.sim, R.E
    rec = {"workload": f"cube{args.strong_n}" if kind == "strong" else "array16_1b", "scene": desc, "scaling": "strong",
           "value": round(sim.cells * K / (ms / 1e3) / 1e6, 1), "unit": "Mcell/s", "ms_per_step": round(ms / K, 5),
           "grid": [sim.nx, sim.ny, sim.nz_glob], "cells": sim.cells, "cells_per_gpu": sim.nx * sim.ny * sim.nz, "n_gpus": world,
           "sample_interval": sim.interval, "gpu_launches": launches, "timed_region": region, "operator_build_s": round(R.build_s, 2),
           "fused_h_to_e": bool(fused), "kernel_ms": [round(v, 4) for v in kms],
           "plan_cells": dict(zip(("plain", "fused_pml", "separate_pml"), E.plan_info())),
           "row_compression": {("E" if w == 0 else "H"): {"row_slots_compressed": v[0], "row_slots_demoted": v[1], "x_vectors": v[2]}
                               for w, v in sim.compression.items()},
           "whole_step_frac": round(BYTES_PER_CELL_STEP * sim.nx * sim.ny * sim.nz / (ms / K / 1e3) / 1e9 / peak, 4),
           "wall_s": round(time.time() - t0, 1)}
    rec.update(extra)
    F._prepared = None
    del R, sim, E, F
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return rec


def run_b200(args):
    import torch
    import torch.distributed as dist
    from b200fdtd import engine as eng_mod

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, W = args.steps, max(3, args.warmup)
    peak, peak_src = measured_peak()

    F, nf, port = build_scene(args, world, args.cells, nrts=10 ** 6)
    R = Runner(F, local, world, args)
    sim, E = R.sim, R.E
    cells = sim.cells
    local_cells = sim.nx * sim.ny * sim.nz

    # ---- resident throughput: K full steps, inputs already in HBM ----
    clocks = ClockSampler(local)
    R.step(W)
    R.barrier()
    if rank == 0:
        clocks.start()
    ms, launches, region = R.resident(K, 0)
    # ---- dominant kernel, timed alone on the same data, same stream ----
    plain_cells, fused_cells, sep_cells = E.plan_info()
    fused_he, kms = R.kernel_times()
    clk = clocks.stop() if rank == 0 else None
    value = cells * K / (ms / 1e3) / 1e6

    # ---- end to end through the public API with HOST buffers ----
    # FDTD.Run(sim_path) on the prepared scene: the operator goes host -> device in the form the host builder produces it
    # (x-vector tables + 32-byte row records + the non-separable rows, pinned host memory), is expanded and verified on the
    # device, K steps run (energy / end-criteria check included), every result comes back to the host (probe series and DFTs,
    # NF2FF face spectra); then port.CalcPort (201 frequencies) and the reference's per-phi nf2ff.CalcNF2FF loop.
    h2d = sim.operator_nbytes(sim.host_op) if getattr(sim, "host_op", None) is not None else 0
    before = None
    if getattr(sim, "host_op", None) is not None:
        big = local_cells > 300e6          # no room for a second copy of the operator: fp64 checksums + sample planes
        if big:
            before = [(float(torch.sum(t, dtype=torch.float64)), t[:, 1::max(1, sim.nz // 7)].clone()) for t in (E.vv, E.vi, E.ii, E.iv)]
        else:
            before = [t.clone() for t in (E.vv, E.vi, E.ii, E.iv)]
        for t in (E.vv, E.vi, E.ii, E.iv):
            t.zero_()
    theta, phis = np.arange(0.0, 181.0, 10.0), np.arange(0.0, 360.0, 45.0)
    fgrid = np.linspace(1e9, 4e9, 201)
    # untimed warm-up of the same call sequence (3 steps): first launches of the reload / far-field kernels load their modules
    F.SetNumberOfTimeSteps(3)
    F.Run(R.path)
    if port is not None:
        port.CalcPort(R.path, fgrid)
    if nf is not None:
        nf.CalcNF2FF(R.path, 2.45e9, theta, phis[:1], center=[0.0, 0.0, 0.8e-3])
    if before is not None:
        for t in (E.vv, E.vi, E.ii, E.iv):
            t.zero_()
    F.SetNumberOfTimeSteps(K)
    R.barrier()
    t_e0 = time.perf_counter()
    F.Run(R.path)
    assert F.sim is sim and sim.timesteps == K
    t_run = time.perf_counter() - t_e0
    if port is not None:
        port.CalcPort(R.path, fgrid)
    t_port = time.perf_counter() - t_e0 - t_run
    t_nf = []
    if nf is not None:
        for ph in phis:
            t1 = time.perf_counter()
            nf.CalcNF2FF(R.path, 2.45e9, theta, np.array([ph]), center=[0.0, 0.0, 0.8e-3])
            t_nf.append(time.perf_counter() - t1)
    R.barrier()
    e2e_s = time.perf_counter() - t_e0
    breakdown = {"run_s": round(t_run, 4), "restart_s": round(getattr(sim, "restart_s", 0.0), 4), "stepping_s": round(sim.wall_s, 4),
                 "collect_s": round(getattr(sim, "collect_s", 0.0), 4), "calcport_s": round(t_port, 4),
                 "calcnf2ff_first_s": round(t_nf[0], 4) if t_nf else None, "calcnf2ff_rest_s": round(sum(t_nf[1:]), 4) if t_nf else None}
    e2e_ok = None
    if before is not None:
        if local_cells > 300e6:
            e2e_ok = [int(float(torch.sum(t, dtype=torch.float64)) != cs) + int((t[:, 1::max(1, sim.nz // 7)] != smp).count_nonzero())
                      for (cs, smp), t in zip(before, (E.vv, E.vi, E.ii, E.iv))]
        else:
            e2e_ok = [sum(int((a[c] != b[c]).count_nonzero()) for c in range(3)) for a, b in zip(before, (E.vv, E.vi, E.ii, E.iv))]
        del before
    outs = [o for o in [E.series, E.probe_dft] + list(E.face_acc) if o is not None]
    d2h = sum(o.numel() * o.element_size() for o in outs)
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = cells * K / e2e_s / 1e6

    line = None
    if rank == 0:
        kname_f = "update_he6_kernel<7> fused H->E launch over the plain region (one launch = both passes)"
        kname_s = "update_e_kernel<4,0,1>/update_h_kernel<4,0,1> plain launch (mean of both passes)"
        roof = roofline_record(fused_he, kms, plain_cells, peak, peak_src, kname_f, kname_s,
                               ("dram_bytes_per_cell_fused_launch", "dram_bytes_per_cell_pass"))
        roof.update({"fused_pml_cells": fused_cells, "separate_pml_cells": sep_cells,
                     "whole_step_frac": round(BYTES_PER_CELL_STEP * local_cells / (ms / K / 1e3) / 1e9 / peak, 4)})
        line = {
            "metric": "FDTD Mcell-updates/s", "value": round(value, 1), "unit": "Mcell/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms / K, 5), "higher_is_better": True, "scaling": "weak" if args.workload == "patch100m" else "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "patch100m" if args.workload == "patch100m" else f"cube{args.n}",
                       "scene": ("2.45 GHz FR-4 patch (reference recipe), PML_8, Gaussian 1-4 GHz, lumped port, V/I probes + DFT, NF2FF DFT"
                                 if args.workload == "patch100m" else "uniform vacuum cube, Mur on 6 faces, centre soft source, one V probe (config 5)"),
                       "grid": [sim.nx, sim.ny, sim.nz_glob], "cells": cells, "cells_per_gpu": local_cells, "pml_cells_rank0": sim.pml_cells,
                       "timestep_s": sim.dt, "sample_interval": sim.interval, "timed_region": region, "parallelism": f"z-slab x{world}",
                       "l2_note": "working set 72 B/cell >> 126 MB L2 (inputs larger than L2, no flush needed)",
                       "operator_build_s": round(R.build_s, 2),
                       "row_compression": {("E" if w == 0 else "H"): {"row_slots_compressed": v[0], "row_slots_demoted": v[1], "x_vectors": v[2]}
                                           for w, v in sim.compression.items()}},
            "roofline": roof,
            "e2e": {"value": round(e2e_value, 1), "unit": "Mcell/s", "h2d_bytes_per_step": int(h2d / K), "d2h_bytes_per_step": int(d2h / K),
                    "what": "public API on the prepared scene: FDTD.Run(sim_path) [compressed operator from pinned host memory -> device, "
                            "expanded + verified on the device, K steps with energy check, probe series/DFT + NF2FF spectra -> host], "
                            "port.CalcPort(201 f), nf2ff.CalcNF2FF per phi (19 theta x 8 phi)",
                    "operator_restored_mismatches": e2e_ok, "seconds": round(e2e_s, 4), "breakdown": breakdown},
            "gpu_launches": int(launches),
            "clocks": clk,
        }
    # ---- second record, N = 1: the seeded per-cell-coefficient operator on the same grid ----
    if world == 1 and not args.no_seeded and args.workload == "patch100m":
        try:
            rec = seeded_record(R, K, W, peak, peak_src)
        except Exception as e:
            rec = {"failed": repr(e)}
        if line is not None:
            line["seeded"] = rec
    F._prepared = None
    del R, sim, E, F, nf, port
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    # ---- latency regime, N = 1: the reference's own ~0.3 M-cell scene end to end ----
    if world == 1 and not args.no_latency and args.workload == "patch100m":
        try:
            rec = latency_record(args, local)
        except Exception as e:
            rec = {"failed": repr(e)}
        if line is not None:
            line["latency"] = rec
    # ---- strong-scaling record (every N, so the driver's N = 1, 2, 4, 8 lines carry the series) and config 3 (N > 1) ----
    for kind, on in (("strong", not args.no_strong), ("config3", world > 1 and not args.no_config3)):
        if not on or args.workload != "patch100m":
            continue
        try:
            rec = sub_workload(kind, args, local, world, K, W, peak, peak_src)
        except Exception as e:
            rec = {"failed": repr(e)}
            if world > 1:
                raise
        if line is not None:
            line[kind] = rec
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
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
