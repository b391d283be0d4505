"""The C-ABI library loads and exports every symbol include/b200fdtd.h declares (no compute calls without a GPU),
and the product fails loudly instead of falling back when CUDA is unavailable."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200fdtd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200fdtd_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from b200fdtd import _lib
    names = _declared()
    assert len(names) >= 20
    L = ctypes.CDLL(_lib.SO_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in b200fdtd.h but not exported by libb200fdtd.so"
    assert sorted(_lib.SYMBOLS) == names, "ctypes binding table and header disagree"
    lib = _lib.lib()
    assert lib.b200fdtd_version() >= 1
    assert lib.b200fdtd_last_error() is not None


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from b200fdtd import B200FDTDError
    from b200fdtd.engine import Engine
    with pytest.raises(B200FDTDError):
        Engine(16, 16, 16)
    # the drop-in Run() must raise too (the reference's try/except turns it into ok=False)
    import scenes
    scenes.use_cuda_engine()
    F, nf, port = scenes.dipole("MUR", cells=(12, 12, 14), nrts=10)
    with pytest.raises(B200FDTDError):
        F.Run(scenes.tmp_sim_path("nofallback"), cleanup=True)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fdtd-solver-antennas_b200")
    for dp, dn, fn in os.walk(pkg):
        for f in fn:
            if f.endswith((".py", ".cu", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "from oracle" not in txt and "import oracle" not in txt and "fdtd_ref" not in txt.replace("oracle/fdtd_ref.c", ""), f
