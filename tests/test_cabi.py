"""The C-ABI shared library loads on a CPU-only box, exports every symbol include/obca.h declares, and refuses to
compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "obca.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(obca_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    import obca_b200
    lib = obca_b200.lib()
    names = _declared()
    assert "obca_parking_solve_batch" in names and "obca_dualmultws_batch" in names and len(names) >= 8
    for n in names:
        assert hasattr(lib, n), n
    assert lib.obca_version() == 200


def test_default_opts_match_reference_call_site():
    """ParkingSignedDist.jl:41-43: tol=1e-5, max_iter=200, min_hessian_perturbation=1e-12."""
    import obca_b200
    o = obca_b200.default_opts()
    assert o.tol == 1e-5 and o.max_iter == 200 and o.dw_min == 1e-12 and o.mu_init == 0.1 and o.retry == 1
    assert o.kappa_eps == 10.0 and o.tau_min == 0.99 and o.gamma_theta == 1e-5


def test_no_cpu_fallback():
    import obca_b200
    from obca_b200 import parking, scenarios
    if obca_b200.lib().obca_device_count() > 0:
        pytest.skip("CUDA device present")
    sc = scenarios.reverse_parking_batch(1, 20, 0)
    with pytest.raises(obca_b200.ObcaError, match="no CUDA device"):
        parking.parking_solve_batch(sc["x0"], sc["xF"], 20, sc["Ts"], sc["L"], sc["ego"], sc["XYbounds"], 3, sc["vOb"],
                                    sc["A"], sc["b"], sc["rx"], sc["ry"], sc["ryaw"], 0, sc["xWS"], sc["uWS"])
    with pytest.raises(obca_b200.ObcaError):
        parking.dualmultws_batch(20, 3, sc["vOb"], sc["A"], sc["b"], sc["rx"], sc["ry"], sc["ryaw"], sc["ego"])


def test_argument_errors():
    import obca_b200
    lib = obca_b200.lib()
    rc = lib.obca_parking_solve_batch(C.c_int(0), C.c_int(80), C.c_int(3), *([None] * 5), C.c_double(0.6), C.c_double(2.7),
                                      *([None] * 9), C.c_int(0), C.c_int(1), None, *([None] * 10))
    assert rc == -1
    assert b"null" in lib.obca_last_error()


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may touch oracle/."""
    for dp, _, files in os.walk(os.path.join(ROOT, "obca_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace(
                    "oracle/ipm_ref.py is the independent restatement", ""), f
