import sys, os, time, cProfile, pstats
sys.path[:0]=['/root/repo','/root/repo/fdtd-solver-antennas_b200','/root/repo/tests']
import numpy as np, torch
import replay, scenes
scenes.use_cuda_engine()
for it in range(2):
    R = replay.replay("trace_single_pml8_q3"); F=R["FDTD"]; F.SetNumberOfTimeSteps(200)
    pr=cProfile.Profile(); pr.enable()
    t=time.time(); F.Run(f"/tmp/pp{it}", cleanup=True); print("run", it, time.time()-t, "prepare", F.sim.prepare_s)
    pr.disable()
    if it==1: pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
