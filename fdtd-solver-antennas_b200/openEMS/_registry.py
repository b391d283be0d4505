"""Results of finished runs, keyed by sim_path (CalcNF2FF / CalcPort receive only the path)."""
import os
import threading

_RESULTS = {}
_LOCK = threading.Lock()


def store(sim_path, res):
    with _LOCK:
        _RESULTS[os.path.abspath(str(sim_path))] = res


def results_for(sim_path, fdtd=None):
    with _LOCK:
        res = _RESULTS.get(os.path.abspath(str(sim_path)))
    if res is None and fdtd is not None and getattr(fdtd, "results", None) is not None:
        res = fdtd.results
    return res
