"""The C-ABI library loads on a machine without a GPU and exports every symbol that
include/uf3b.h declares; host-only entry points behave as documented."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from uf3_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "uf3b.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(uf3b_[a-z0-9_]+)\s*\(", text)))


def test_header_functions_are_exported_and_bound():
    lib = _native.lib()
    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/uf3b.h but not exported"
        assert name in _native.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.uf3b_abi_version() == 1


def test_descriptor_layout_matches_header():
    text = open(os.path.join(ROOT, "include", "uf3b.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} uf3b_basis_desc;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 2 if decl.startswith("const") else 1)[-1]
        fields += [n.strip().lstrip("*") for n in names.split(",")]
    assert fields == [name for name, _ in _native.BasisDesc._fields_]


def test_host_spline_probe_matches_reference_known_values():
    """tests/test_bsplines.py:529-547 of the reference: clamped knots [0,0,0,0,1,1,1,1]."""
    lib = _native.lib()
    knots = np.array([0, 0, 0, 0, 1, 1, 1, 1], dtype=np.float64)
    v, dv = (C.c_double * 4)(), (C.c_double * 4)()
    kp = knots.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.uf3b_host_eval_basis(kp, 8, 1e-10, v, dv) == 0
    assert np.allclose(v[:], [1, 0, 0, 0], atol=1e-8)
    assert lib.uf3b_host_eval_basis(kp, 8, 0.5, v, dv) == 0
    assert np.allclose(v[:], [0.125, 0.375, 0.375, 0.125])
    assert lib.uf3b_host_eval_basis(kp, 8, 1 - 1e-10, v, dv) == 0
    assert np.allclose(v[:], [0, 0, 0, 1], atol=1e-8)
    # outside (knots[3], knots[-4]]: no contribution
    assert lib.uf3b_host_eval_basis(kp, 8, 0.0, v, dv) == -1
    assert lib.uf3b_host_eval_basis(kp, 8, 1.5, v, dv) == -1
    assert lib.uf3b_host_eval_basis(kp, 8, 1.0, v, dv) == 0


def test_host_spline_probe_matches_scipy_and_oracle():
    from scipy import interpolate
    from oracle import uf3_oracle as orc
    lib = _native.lib()
    rng = np.random.default_rng(3)
    for knots in (np.concatenate([[1.5] * 3, np.linspace(1.5, 7.0, 13), [7.0] * 3]),
                  np.concatenate([[2.0] * 3, np.linspace(4.0, 36.0, 21) ** 0.5, [6.0] * 3])):
        knots = np.ascontiguousarray(knots)
        kp = knots.ctypes.data_as(C.POINTER(C.c_double))
        elements = [interpolate.BSpline.basis_element(knots[i:i + 5], extrapolate=False)
                    for i in range(len(knots) - 4)]
        points = np.concatenate([rng.uniform(knots[0], knots[-1], 300), knots[3:-3]])
        v, dv = (C.c_double * 4)(), (C.c_double * 4)()
        for r in points:
            idx = lib.uf3b_host_eval_basis(kp, len(knots), float(r), v, dv)
            want_idx, want_v, want_dv = orc.eval_basis(knots, float(r))
            assert idx == want_idx == (np.searchsorted(knots, r, side="left") - 4
                                       if knots[3] < r <= knots[-4] else -1)
            if idx < 0:
                continue
            assert np.allclose(v[:], want_v, rtol=0, atol=1e-14)
            assert np.allclose(dv[:], want_dv, rtol=0, atol=1e-13 * max(1.0, np.abs(want_dv).max()))
            if r < knots[-1]:    # scipy's half-open intervals drop the very last knot
                for q in range(4):
                    ref = np.nan_to_num(elements[idx + q](r))
                    assert abs(v[q] - ref) < 1e-14


def test_bad_arguments_report_an_error():
    lib = _native.lib()
    v, dv = (C.c_double * 4)(), (C.c_double * 4)()
    assert lib.uf3b_host_eval_basis(None, 8, 0.5, v, dv) == _native.ERR_INVALID
    assert b"knot" in lib.uf3b_last_error()
    with pytest.raises(_native.UF3BError):
        _native.check(_native.ERR_INVALID)
