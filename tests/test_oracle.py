"""The oracle against everything that pins it: frozen known answers for the reference's test
programs (tests/golden/known_answers.json, partly hand-derived), the independent C++/OpenMP
restatement, and analytic invariants."""
import json
import os

import numpy as np
import pytest

from oracle import reference_cpp as rc
from oracle import reference_numpy as rn

from conftest import GOLDEN, HALO, all_programs, program_path, random_inputs

with open(os.path.join(GOLDEN, "known_answers.json")) as _f:
    KNOWN = json.load(_f)


@pytest.mark.parametrize("name", sorted(KNOWN))
def test_numpy_oracle_matches_known_answers(name):
    res = rn.run_reference(program_path(name))
    for field, rec in KNOWN[name].items():
        got = res[field]
        assert got.dtype.name == rec["dtype"] and list(got.shape) == rec["shape"]
        tol = 1e-6 if rec["dtype"] == "float32" else 1e-13
        assert abs(float(got.sum(dtype=np.float64)) - rec["sum"]) <= tol * max(1.0, abs(rec["sum"]))
        if "values" in rec:
            np.testing.assert_allclose(got, np.array(rec["values"]), rtol=tol, atol=0)
        for point, value in rec.get("points", []):
            assert abs(float(got[tuple(point)]) - value) <= tol * max(1.0, abs(value))


@pytest.mark.parametrize("name", all_programs())
def test_cpp_oracle_matches_numpy_oracle(name):
    path = program_path(name)
    inputs = random_inputs(name)
    a = rn.run_reference(path, inputs)
    b = rc.run_reference_cpp(path, inputs)
    for field in a:
        tol = 2e-6 if a[field].dtype == np.float32 else 1e-13   # -ffast-math reassociation only
        h = HALO.get(name, 0)
        assert rn.max_relative_error(rn.trim_halo(a[field], h), rn.trim_halo(b[field], h)) <= tol, field


def test_jacobi_interior_invariant():
    # constant input c, constant boundary 0: a cell farther than n steps from every face holds
    # c * (6 * 0.16666666)^n, rounded to float32 after every step
    res = rn.run_reference(program_path("ref_jacobi3d_32x32x32_8itr_8vec"))["b7"]
    v = np.float32(1.0)
    for _ in range(8):
        v = np.float32(np.float64(0.16666666) * (np.float64(v) * 6.0))
    assert res[16, 16, 16] == v
    assert res[8, 20, 12] == v
    # symmetry under axis permutation and reflection
    assert np.array_equal(res, res.transpose(2, 0, 1))
    assert np.array_equal(res, res[::-1, :, :])


def test_shrink_interior_independent_of_junk():
    # cells at least (sum of extents) away from the border must not see -100000
    path = program_path("jacobi2d_96x128_6itr_shrink_f64")
    inputs = random_inputs("jacobi2d_96x128_6itr_shrink_f64")
    res = rn.run_reference(path, inputs)["b5"]
    inner = rn.trim_halo(res, 6)
    assert np.all(np.abs(inner) <= 1.0 + 1e-12)
    assert np.any(np.abs(res) > 10.0)      # the halo does carry junk


def test_lower_dimensional_inputs_take_leading_elements():
    # run_program allocates every array input at the full program shape (helper.py:162-217)
    path = program_path("ref_varying_dimensionality")
    full = {"in1d": np.full((8, 16, 32), 0.2, np.float32)}
    a = rn.run_reference(path, full)["out"]
    b = rn.run_reference(path)["out"]
    assert np.array_equal(a, b)
    assert abs(float(a[0, 0, 0]) - 2.7) < 1e-6 and abs(float(a[7, 15, 31]) - 4.0) < 1e-6


def test_linearity_of_jacobi():
    path = program_path("jacobi3d_16x24x32_5itr_const1")
    prog = rn.load_program(path)
    for entry in prog["program"].values():
        for bc in entry["boundary_conditions"].values():
            bc["value"] = 0.0
    rng = np.random.default_rng(3)
    x = rng.uniform(size=(16, 24, 32)).astype(np.float32)
    y = rng.uniform(size=(16, 24, 32)).astype(np.float32)
    fx = rn.run_reference(prog, {"a": x})["b4"].astype(np.float64)
    fy = rn.run_reference(prog, {"a": y})["b4"].astype(np.float64)
    fxy = rn.run_reference(prog, {"a": x + y})["b4"].astype(np.float64)
    assert np.max(np.abs(fxy - fx - fy)) < 5e-6
