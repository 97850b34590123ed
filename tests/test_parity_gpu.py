"""Parity tests proper: the CUDA path (through the C-ABI) against the oracle on the same seeded
inputs.  Tolerances are the north-star ones: max relative error 1e-5 for float32 results and 1e-12
for float64 results, the -halo border excluded for shrink boundaries."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, HALO, all_programs, program_path, random_inputs

pytestmark = pytest.mark.gpu

TOL = {"float32": 1e-5, "float64": 1e-12}


@pytest.fixture(scope="module")
def gpu(native_lib):
    from stencilflow_b200 import runtime
    return runtime.Runtime.get()


def _run_cuda(name, inputs, plan_options=None):
    from stencilflow_b200.cuda_program import CudaProgram
    from oracle import reference_numpy as rn
    prog = CudaProgram(program_path(name), plan_options=plan_options)
    info = rn.ProgramInfo(rn.load_program(program_path(name)))
    outs = {o: np.zeros(info.shape, dtype=info.field_type(o)) for o in info.outputs}
    args = {}
    for k, v in inputs.items():
        args[k + "_host" if getattr(v, "ndim", 0) > 0 else k] = v
    for k, v in outs.items():
        args[k + "_host"] = v
    prog(**args)
    prog.close()
    return outs, prog


def _check(name, got, expected):
    from oracle import reference_numpy as rn
    h = HALO.get(name, 0)
    for field, ref in expected.items():
        err = rn.max_relative_error(rn.trim_halo(ref, h), rn.trim_halo(got[field], h))
        assert err <= TOL[ref.dtype.name], "{}:{} max rel err {}".format(name, field, err)


@pytest.mark.parametrize("name", all_programs())
def test_program_matches_oracle_general(gpu, name):
    """Every program, every operator as its own launch (fusion off)."""
    from oracle import reference_numpy as rn
    from stencilflow_b200.planner import PlanOptions
    inputs = random_inputs(name)
    expected = rn.run_reference(program_path(name), inputs)
    got, _ = _run_cuda(name, inputs, PlanOptions(fuse=False))
    _check(name, got, expected)


@pytest.mark.parametrize("name", all_programs())
def test_program_matches_oracle_planned(gpu, name):
    """Every program with the planner's default choice (fused passes where streamable)."""
    from oracle import reference_numpy as rn
    inputs = random_inputs(name, seed=11)
    expected = rn.run_reference(program_path(name), inputs)
    got, _ = _run_cuda(name, inputs)
    _check(name, got, expected)


@pytest.mark.parametrize("name", [n for n in all_programs() if n.startswith("ref_")])
def test_reference_programs_match_known_answers(gpu, name):
    """The reference's own programs with their own inputs against the frozen known answers."""
    from oracle import reference_numpy as rn
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        known = json.load(f)[name]
    inputs = rn.materialize_inputs(rn.load_program(program_path(name)))
    got, _ = _run_cuda(name, inputs)
    for field, rec in known.items():
        tol = TOL[rec["dtype"]]
        arr = got[field]
        assert abs(float(arr.sum(dtype=np.float64)) - rec["sum"]) <= 10 * tol * max(1.0, abs(rec["sum"]))
        if "values" in rec:
            assert rn.max_relative_error(np.array(rec["values"], dtype=arr.dtype), arr) <= tol
        for point, value in rec.get("points", []):
            assert abs(float(arr[tuple(point)]) - value) <= tol * max(1.0, abs(value))


@pytest.mark.parametrize("name,halo", [("ref_jacobi3d_32x32x32_8itr_8vec", 0), ("ref_varying_dimensionality", 0),
                                       ("ref_simulator11", 0), ("hdiff_24x28x16", 2),
                                       ("jacobi2d_96x128_6itr_shrink_f64", 6)])
def test_run_program_compare_to_reference(gpu, name, halo, tmp_path, monkeypatch):
    """`run_program.py prog.json cuda -compare-to-reference` end to end (drop-in driver)."""
    from stencilflow_b200.run_program import run_program
    monkeypatch.chdir(tmp_path)
    ret = run_program(program_path(name), "cuda", compare_to_reference=True, halo=halo, log_level=0,
                      input_directory=os.path.dirname(program_path(name)))
    assert ret == 0
    base = os.path.splitext(os.path.basename(program_path(name)))[0].replace(".", "_")
    assert os.path.isdir(tmp_path / "results" / base / "reference")


def test_mismatch_is_detected(gpu, tmp_path, monkeypatch):
    from stencilflow_b200 import run_program as rp
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(rp, "tolerance_for", lambda dtype: -1.0)
    with pytest.raises(ValueError, match="Result mismatch"):
        rp.run_program(program_path("ref_simulator3"), "cuda", compare_to_reference=True, log_level=0)


def test_generate_input_and_repetitions(gpu, tmp_path, monkeypatch):
    from stencilflow_b200.run_program import run_program
    monkeypatch.chdir(tmp_path)
    assert run_program(program_path("ref_jacobi2d_128x128"), "cuda", compare_to_reference=True,
                       generate_input=True, repetitions=3, log_level=0) == 0
    assert run_program(program_path("ref_varying_dimensionality"), "cuda", compare_to_reference=True,
                       specialize_scalars=True, log_level=0) == 0
    assert run_program(program_path("ref_jacobi2d_128x128"), "cuda", compare_to_reference=True,
                       synthetic_reads=0.75, log_level=0) == 0


def test_device_utilities(gpu):
    """fill_hash is bit-identical to its host mirror; checksum/compare agree with numpy."""
    from stencilflow_b200 import synthetic
    n = 1 << 20
    for dt in (np.float32, np.float64):
        d = gpu.malloc(n * np.dtype(dt).itemsize)
        gpu.fill_hash(d, n, dt, seed=1234, lo=-1.0, hi=3.0, index_offset=5)
        host = np.empty(n, dtype=dt)
        gpu.d2h(host, d)
        gpu.stream_synchronize()
        mirror = synthetic.fill_hash((n,), dt, 1234, -1.0, 3.0, 5)
        assert np.array_equal(host, mirror)
        s, bits = gpu.checksum(d, n, dt)
        assert abs(s - float(host.sum(dtype=np.float64))) <= 1e-9 * n
        words = host.view(np.uint32 if dt == np.float32 else np.uint64).astype(np.uint64)
        assert bits == int(words.sum(dtype=np.uint64))
        d2 = gpu.malloc(n * np.dtype(dt).itemsize)
        other = host.copy()
        other[12345] *= dt(1.001)
        gpu.h2d(d2, other)
        m, bad = gpu.compare(d, d2, n, dt, 1e-5)
        assert bad == 1 and abs(m - 1e-3) < 2e-4
        gpu.free(d)
        gpu.free(d2)
    # indices beyond 2^32 use the high word
    a = synthetic.fill_hash((4,), np.float32, 1, index_offset=(1 << 33))
    b = synthetic.fill_hash((4,), np.float32, 1, index_offset=0)
    assert not np.array_equal(a, b)


def test_size_independent_properties_at_scale(gpu):
    """256^3 x 8 chained Jacobi on hash-generated input, checked by properties that do not need the
    oracle to hold nine 256^3 arrays: linearity and agreement of fused vs unfused execution."""
    from stencilflow_b200.cuda_program import CudaProgram
    from stencilflow_b200.planner import PlanOptions
    from stencilflow_b200 import programs, synthetic
    path = programs.write_program(programs.jacobi3d_chain([256, 256, 256], 8), "jacobi3d_256_8itr")
    n = 256 ** 3
    results = []
    for opts in (PlanOptions(fuse=False), None):
        p = CudaProgram(path, plan_options=opts)
        p.rt.fill_hash(p.buffers["a"].dptr, n, np.float32, seed=1234)
        p.execute()
        results.append(p.download("b7"))
        p.close()
    from oracle import reference_numpy as rn
    assert rn.max_relative_error(results[0], results[1]) <= 1e-5
    # oracle on a 64^3 corner problem with the same generator: exact same program at reduced size
    small = programs.write_program(programs.jacobi3d_chain([64, 64, 64], 8), "jacobi3d_64_8itr")
    a = synthetic.fill_hash((64, 64, 64), np.float32, 1234)
    expected = rn.run_reference(small, {"a": a})["b7"]
    p = CudaProgram(small)
    out = np.zeros((64, 64, 64), np.float32)
    p(a_host=a, b7_host=out)
    p.close()
    assert rn.max_relative_error(expected, out) <= 1e-5
