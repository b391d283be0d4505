import sys, os, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,"fdtd-solver-antennas_b200")]
import torch
from b200fdtd import scenes
from b200fdtd.simulation import Simulation
F, nf, port = scenes.patch_scene(target_cells=100e6, boundary="PML_8", f0=2.5e9, fc=1.5e9, nrts=10**6, end_criteria=1e-12, nf2ff_freqs=[2.45e9])
S=F._setup(); sim=Simulation(S, device=0, nf2ff_freqs=F.nf2ff_freqs, probe_freqs=S.probe_freqs); sim.prepare(); E=sim.engine
E.run(105, use_graph=True); torch.cuda.synchronize()
op=sim.export_operator(pin=True)
for rep in range(2):
    for t in (E.vv,E.vi,E.ii,E.iv): t.zero_()
    torch.cuda.synchronize()
    t0=time.perf_counter(); sim.load_operator(op); torch.cuda.synchronize(); t1=time.perf_counter()
    E.run(100, use_graph=True); torch.cuda.synchronize(); t2=time.perf_counter()
    outs=[E.series,E.probe_dft]+list(E.face_acc); res=[o.to("cpu") for o in outs]; torch.cuda.synchronize(); t3=time.perf_counter()
    print(f"load {1e3*(t1-t0):.1f} ms, run100 {1e3*(t2-t1):.1f} ms, d2h {1e3*(t3-t2):.1f} ms ({sum(o.numel()*o.element_size() for o in res)/1e6:.1f} MB)")
# breakdown of load
import cProfile, pstats
for t in (E.vv,E.vi,E.ii,E.iv): t.zero_()
torch.cuda.synchronize()
pr=cProfile.Profile(); pr.enable(); sim.load_operator(op); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
