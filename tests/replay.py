"""Replay a recorded reference call trace (tests/golden/trace_*.json) against the CSXCAD/openEMS shim."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _dec(v, objs):
    if isinstance(v, dict):
        if "ref" in v and len(v) == 1:
            return objs[v["ref"]]
        if "nd" in v and len(v) == 1:
            return np.asarray(v["nd"])
        return {k: _dec(x, objs) for k, x in v.items()}
    if isinstance(v, list):
        return [_dec(x, objs) for x in v]
    return v


def replay(case):
    """returns dict(FDTD=..., nf=..., theta=, phi=, nf_center=, objs=)"""
    from CSXCAD import ContinuousStructure
    from openEMS import openEMS
    T = json.load(open(os.path.join(GOLDEN, case + ".json")))
    objs = {}
    for e in T["log"]:
        if "new" in e:
            cls = {"ContinuousStructure": ContinuousStructure, "openEMS": openEMS}[e["new"]]
            objs[e["id"]] = cls(*_dec(e["args"], objs), **_dec(e["kw"], objs))
        else:
            fn = getattr(objs[e["obj"]], e["call"])
            objs[e["ret"]] = fn(*_dec(e["args"], objs), **_dec(e["kw"], objs))
    return dict(FDTD=objs[T["fdtd"]], nf=objs[T["nf"]], theta=np.asarray(T["theta"]), phi=np.asarray(T["phi"]),
                nf_center=np.asarray(T["nf_center"]), objs=objs, trace=T)


def reference_postprocess(nf2ff, sim_path, f_res, theta, phi, nf_center):
    """the reference's far-field loop and dBi conversion, restated for tests
    (antenna_sim/solver_fdtd_openems_microstrip_3d.py:221-248): one CalcNF2FF per phi, E/Emax and Dmax of the first call"""
    E_stack, Dmax = [], None
    for ph in phi:
        res = nf2ff.CalcNF2FF(sim_path, f_res, theta, np.array([ph]), center=nf_center)
        E_stack.append(np.squeeze(np.asarray(res.E_norm[0])).reshape(-1))
        if Dmax is None:
            Dmax = float(np.asarray(res.Dmax)[0])
    E = np.stack(E_stack, axis=1)
    return 20.0 * np.log10(E / E.max() + 1e-16) + 10.0 * np.log10(Dmax), Dmax


def s11_db(port, sim_path, f0):
    """the reference's S11 block (antenna_sim/solver_fdtd_openems_microstrip.py:406-416)"""
    f = np.linspace(max(1e9, f0 * 0.7), f0 * 1.3, 201)
    port.CalcPort(sim_path, f)
    s11 = port.uf_ref / port.uf_inc
    return f, 20.0 * np.log10(np.abs(s11) + 1e-16), port.uf_tot / port.if_tot
