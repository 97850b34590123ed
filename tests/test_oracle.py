"""The oracle against everything that pins it: OUTPUTS OF THE REFERENCE ITSELF (its dataflow simulator
run in the build container, tests/golden/reference_sim.npz, generator make_reference_sim_golden.py),
frozen known answers for the reference's test programs (tests/golden/known_answers.json, partly
hand-derived), the independent C++/OpenMP restatement, and analytic invariants."""
import json
import os

import numpy as np
import pytest

from oracle import reference_cpp as rc
from oracle import reference_numpy as rn

from conftest import GOLDEN, HALO, all_programs, program_path, random_inputs

with open(os.path.join(GOLDEN, "known_answers.json")) as _f:
    KNOWN = json.load(_f)


with open(os.path.join(GOLDEN, "reference_sim.json")) as _f:
    REFERENCE_SIM = json.load(_f)


def reference_sim_case(case):
    """(program dict, inputs, {output: array the reference's simulator produced})"""
    sys_path_golden()
    import make_reference_sim_golden as gen
    rec = REFERENCE_SIM[case]
    prog = gen.case_program(rec["program"])
    inputs = gen.case_inputs(prog, rec["seed"], rec.get("ranges"))
    with np.load(os.path.join(GOLDEN, "reference_sim.npz")) as z:
        expected = {o: z[case + "/" + o] for o in rec["outputs"]}
    return prog, inputs, expected


def sys_path_golden():
    import sys
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)


@pytest.mark.parametrize("case", sorted(REFERENCE_SIM))
def test_oracles_match_reference_simulator(case):
    """Both restatements against what the reference's own code computed for the same program and
    inputs.  The simulator evaluates every operator in Python floats (double) and rounds the result
    to the operator's data type (kernel.py:716-718), the oracle evaluates in the data type itself:
    for float32 that is a few ulp per operator, far inside the 1e-5 of arrays_are_equal."""
    prog, inputs, expected = reference_sim_case(case)
    got_np = rn.run_reference(prog, inputs)
    got_cpp = rc.run_reference_cpp(prog, inputs)
    # cases the simulator ran in an equivalent form (shrink as the constant -100000, 2-D embedded in 3-D,
    # statements inlined; make_reference_sim_golden.py): the ORIGINAL program is run here, interiors compared
    h = REFERENCE_SIM[case].get("halo", 0)
    for field, ref in expected.items():
        tol = 2e-6 if ref.dtype == np.float32 else 1e-13
        assert got_np[field].dtype == ref.dtype and got_np[field].shape == ref.shape
        assert rn.max_relative_error(rn.trim_halo(ref, h), rn.trim_halo(got_np[field], h)) <= tol, (case, field)
        assert rn.max_relative_error(rn.trim_halo(ref, h), rn.trim_halo(got_cpp[field], h)) <= 2 * tol, (case, field)


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
