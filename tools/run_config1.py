"""Config 1 end to end (BASELINE.json configs[0]): the reference's own single-patch scene (recorded call trace of the
unmodified prepare function, GUI defaults: PML_8, mesh quality 3, NrTS 30000, EndCriteria 1e-4), full run on the CUDA
engine and on the CPU oracle; reports wall time, S11, resonance, directivity and the CUDA-vs-oracle deltas."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "fdtd-solver-antennas_b200")]
import numpy as np  # noqa: E402

import replay  # noqa: E402
import scenes  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "trace_single_pml8_q3"
nrts = int(sys.argv[2]) if len(sys.argv) > 2 else None
out = {}
for eng in ("cuda", "oracle"):
    (scenes.use_oracle_engine(threads=os.cpu_count() or 4) if eng == "oracle" else scenes.use_cuda_engine())
    R = replay.replay(case)
    F, nf = R["FDTD"], R["nf"]
    if nrts:
        F.SetNumberOfTimeSteps(nrts)
    path = scenes.tmp_sim_path(f"cfg1_{eng}")
    t0 = time.time()
    F.Run(path, cleanup=True, verbose=0)
    t_run = time.time() - t0
    f0 = F.exc[1]
    t1 = time.time()
    dbi, dmax = replay.reference_postprocess(nf, path, f0, R["theta"], R["phi"], R["nf_center"])
    t_ff = time.time() - t1
    f, s11, zin = replay.s11_db(F.ports[0], path, f0)
    sim = F.sim
    out[eng] = dict(run_s=t_run, farfield_s=t_ff, prepare_s=sim.prepare_s, step_loop_s=sim.wall_s, timesteps=sim.timesteps,
                    stop=sim.stop_reason, cells=sim.cells, grid=[sim.nx, sim.ny, sim.nz_glob], dt=sim.dt,
                    mcells_per_s=sim.cells * sim.timesteps / sim.wall_s / 1e6,
                    s11_min_db=float(s11.min()), f_res_ghz=float(f[s11.argmin()] / 1e9), dmax_dbi=float(10 * np.log10(dmax)),
                    pattern_max_dbi=float(dbi.max()), zin_at_f0=[float(np.real(zin[len(zin) // 2])), float(np.imag(zin[len(zin) // 2]))])
    out[eng + "_arrays"] = dict(s11=s11, dbi=dbi)
    print(eng, json.dumps(out[eng]), flush=True)
c, o = out["cuda_arrays"], out["oracle_arrays"]
sel = o["dbi"] > o["dbi"].max() - 30
delta = dict(s11_max_abs_db=float(np.abs(c["s11"] - o["s11"]).max()),
             pattern_max_abs_db=float(np.abs(c["dbi"][sel] - o["dbi"][sel]).max()),
             dmax_db=float(abs(out["cuda"]["dmax_dbi"] - out["oracle"]["dmax_dbi"])),
             f_res_rel=float(abs(out["cuda"]["f_res_ghz"] - out["oracle"]["f_res_ghz"]) / out["oracle"]["f_res_ghz"]),
             speedup_run=out["oracle"]["run_s"] / out["cuda"]["run_s"], oracle_threads=os.cpu_count())
print("DELTA", json.dumps(delta))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dict(case=case, cuda=out["cuda"], oracle=out["oracle"], delta=delta), open(os.path.join(ROOT, "gpurun_out", f"config1_{case}.json"), "w"), indent=1)
