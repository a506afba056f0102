"""LAMMPS hand-off texts against the files the reference's own
functions wrote for its shipped models (oracle/make_golden_lammps.py)."""
import os
import re

import numpy as np

import golden_util as gu
from uf3_b200 import lammps, least_squares as ls

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _model(case_name):
    case = gu.Case(case_name)
    model = ls.WeightedLinearModel(case.basis())
    model.coefficients = np.array(case["coefficients"])
    return model


def _without_stamp(text):
    return re.sub(r"DATE: \S+ \S+ AUTHOR", "DATE: - AUTHOR", text)


def test_uf3_potential_file_matches_the_reference_text():
    for tag, case in (("W23", "calc_syn_w54_model23"), ("NeXe", "calc_syn_nexe64_pair")):
        model = _model(case)
        want = open(os.path.join(GOLDEN, f"lammps_{tag}.uf3.txt")).read()
        got = lammps.uf3_lammps_pot_text(model, "nk", author="golden", lammps_units="metal",
                                         legacy_trim_field=True)
        assert _without_stamp(got) == _without_stamp(want)
        # default: the block's own trims as two integers, everything else unchanged
        fixed = lammps.uf3_lammps_pot_text(model, "nk", author="golden", lammps_units="metal")
        for a, b in zip(_without_stamp(fixed).splitlines(), _without_stamp(want).splitlines()):
            if a.startswith(("2B", "3B")):
                degree = int(a[0])
                head, tail = a.split()[:degree + 1], a.split()[degree + 1:]
                assert head == b.split()[:degree + 1] and tail[2] == "nk"
                assert [int(tail[0]), int(tail[1])] == [model.bspline_config.leading_trim[degree],
                                                        model.bspline_config.trailing_trim[degree]]
            else:
                assert a == b


def test_writer_keeps_the_reference_signature(tmp_path):
    model = _model("calc_syn_w54_model23")
    path = lammps.write_uf3_lammps_pot_files(model.bspline_config.chemical_system, model, "nk", str(tmp_path / "pots"),
                                             "W.uf3", "me", "metal")
    assert os.path.isfile(path) and open(path).read().startswith("#UF3 POT UNITS: metal")
    assert lammps.lammps_input_lines(model, "pots", "W.uf3").splitlines()[0] == "pair_style\tuf3 3 1"
    try:
        lammps.uf3_lammps_pot_text(model, "xx")
    except ValueError:
        pass
    else:
        raise AssertionError("an unknown knot spacing type must be rejected")


def test_tabulated_pair_potential_matches_the_reference_text():
    model = _model("calc_syn_w54_model23")
    basis = model.bspline_config
    pair = basis.interactions_map[2][0]
    sizes, starts = basis.get_interaction_partitions()[:2]
    got = lammps.export_tabulated_potential(basis.knots_map[pair],
                                            model.coefficients[starts[pair]:starts[pair] + sizes[pair]], pair,
                                            grid=200, contributor="golden", rounding=8)
    want = open(os.path.join(GOLDEN, "lammps_W23_pair_table.txt")).read()
    got_lines, want_lines = got.splitlines(), want.splitlines()
    assert got_lines[1:6] == want_lines[1:6] and len(got_lines) == len(want_lines) == 206
    for a, b in zip(got_lines[6:], want_lines[6:]):
        ia, ra, ea, fa = a.split()
        ib, rb, eb, fb = b.split()
        assert ia == ib and ra == rb
        assert abs(float(ea) - float(eb)) <= 2e-8 * max(1.0, abs(float(eb)))
        assert abs(float(fa) - float(fb)) <= 2e-8 * max(1.0, abs(float(fb)))
