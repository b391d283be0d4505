"""Results of finished runs, keyed by sim_path (CalcNF2FF / CalcPort receive only the path)."""
import os
import threading

_RESULTS = {}
_LOCK = threading.Lock()


MAX_KEPT = 4        # a result keeps device memory alive (DFT accumulators, stored NF2FF face samples): only the latest runs


def store(sim_path, res):
    with _LOCK:
        key = os.path.abspath(str(sim_path))
        _RESULTS.pop(key, None)
        _RESULTS[key] = res
        while len(_RESULTS) > MAX_KEPT:
            _RESULTS.pop(next(iter(_RESULTS)))


def results_for(sim_path, fdtd=None):
    with _LOCK:
        res = _RESULTS.get(os.path.abspath(str(sim_path)))
    if res is None and fdtd is not None and getattr(fdtd, "results", None) is not None:
        res = fdtd.results
    return res
