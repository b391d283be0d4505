import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "fdtd-solver-antennas_b200")]
import torch, numpy as np
from b200fdtd import scenes as ps
from b200fdtd.simulation import Simulation
F, nf, port = ps.patch_scene(target_cells=6e6, boundary="PML_8", f0=2.5e9, fc=1.5e9, nrts=1000, end_criteria=1e-12, nf2ff_freqs=[2.45e9])
S = F._setup()
sim = Simulation(S, device=0, nf2ff_freqs=F.nf2ff_freqs, probe_freqs=S.probe_freqs); sim.prepare()
E = sim.engine
print('build dev', sim.build_device, 'compression', sim.compression)
op = sim.export_operator(pin=True)
print('bytes', sim.operator_nbytes(op), [[len(i) for i in op[w]['full_idx']] for w in (0,1)])
before = [t.clone() for t in (E.vv, E.vi, E.ii, E.iv)]
for t in (E.vv, E.vi, E.ii, E.iv): t.zero_()
torch.cuda.synchronize()
res = sim.load_operator(op)
torch.cuda.synchronize()
print('reload verify', res)
for n, a, b in zip(('vv','vi','ii','iv'), before, (E.vv, E.vi, E.ii, E.iv)):
    d = (a != b)
    print(n, int(d.sum()), a.numel(), float(b.abs().max()), float(a.abs().max()))
    if d.any():
        idx = d.nonzero()[0].tolist(); print('   first', idx, a[tuple(idx)].item(), b[tuple(idx)].item())
