"""The dataflow simulator (``-run-simulation``): against the outputs AND cycle counts of the
reference's own simulator (tests/golden/reference_sim.*, produced by running the reference's code,
see make_reference_sim_golden.py), against the oracle for everything outside that simulator's envelope
(2-D and 1-D programs, lower-dimensional and 0-D inputs, shrink/copy boundaries, and/or), and as a
check of the delay-buffer analysis (occupancies stay within the analysed capacities; a join the
analysis under-sizes is reported as a dead-lock, as the reference's simulator reports an overflow)."""
import inspect
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, HALO, program_path

from oracle import reference_numpy as rn
from stencilflow_b200 import helper
from stencilflow_b200.kernel_chain_graph import KernelChainGraph
from stencilflow_b200.simulator import SimulationDeadlock, Simulator

if GOLDEN not in sys.path:
    sys.path.insert(0, GOLDEN)
import make_reference_sim_golden as gen  # noqa: E402

with open(os.path.join(GOLDEN, "reference_sim.json")) as _f:
    REFERENCE_SIM = json.load(_f)


def _simulate(path):
    chain = KernelChainGraph(path)
    sim = Simulator(program_name=os.path.basename(path)[:-5], program_description=helper.parse_json(path),
                    input_nodes=chain.input_nodes, kernel_nodes=chain.kernel_nodes,
                    output_nodes=chain.output_nodes, dimensions=chain.dimensions, write_output=False, log_level=0)
    sim.simulate()
    return sim


# the 32^3 x 8 chains take ~10 s each in pure Python: one of them is enough
# (cases the reference ran in an equivalent, transformed form -- shrink as a constant, 2-D embedded in 3-D,
# statements inlined -- have other cycle counts than the original program: they pin the oracle, not this model)
FAST_CASES = [c for c in sorted(REFERENCE_SIM)
              if c != "ref_jacobi3d_32x32x32_8itr_8vec_rand" and "transforms" not in REFERENCE_SIM[c]]


@pytest.mark.parametrize("case", FAST_CASES)
def test_matches_reference_simulator_results_and_cycles(case, tmp_path):
    rec = REFERENCE_SIM[case]
    prog = gen.case_program(rec["program"])
    for k, v in gen.case_inputs(prog, rec["seed"]).items():
        prog["inputs"][k]["data"] = [float(x) for x in np.ravel(v)]
    path = str(tmp_path / (case + ".json"))
    with open(path, "w") as f:
        json.dump(prog, f)
    sim = _simulate(path)
    with np.load(os.path.join(GOLDEN, "reference_sim.npz")) as z:
        for field in rec["outputs"]:
            ref = z[case + "/" + field]
            got = sim.get_result()[field].reshape(ref.shape)
            assert got.dtype == ref.dtype
            # both evaluate in Python floats and round once per operator
            assert rn.max_relative_error(ref, got) <= (5e-7 if ref.dtype == np.float32 else 1e-15)
    # same machine, same pace: the reference needed rec["cycles"] cycles
    assert abs(sim.cycles - rec["cycles"]) <= max(4, 0.002 * rec["cycles"]), (sim.cycles, rec["cycles"])
    for name, (capacity, used) in sim.channel_usage().items():
        assert used <= capacity, name


@pytest.mark.parametrize("name", [
    "ref_jacobi2d_128x128", "ref_varying_dimensionality", "ref_simulator", "ref_simulator7", "ref_simulator11",
    "ref_simple_input_delay_buf", "smooth1d_256", "lowdim2d_32x64", "lowdim3d_20x24x48_3st_f32", "math_ops_8x8x8",
    "jacobi3d_12x12x16_3itr_copy", "jacobi2d_96x128_6itr_shrink_f64", "hdiff_16x20x8_f64",
    "synth_hotspot2d_48x64_4st_f64", "synth_diffusion_10x12x16_4st", "jacobi2d_64x256_6itr_w1d_const_f32",
    "hdiff_const_10x12x8_f64", "multistmt3d_6x8x10_f64"])
def test_matches_oracle_outside_the_reference_envelope(name):
    path = program_path(name)
    sim = _simulate(path)
    expected = rn.run_reference(path)
    h = HALO.get(name, 0)
    for field, ref in expected.items():
        got = sim.get_result()[field].reshape(ref.shape)
        tol = 1e-6 if ref.dtype == np.float32 else 1e-13
        assert rn.max_relative_error(rn.trim_halo(ref, h), rn.trim_halo(got, h)) <= tol, field
    assert sim.cycles >= sim.total
    assert sim.all_done()


@pytest.mark.parametrize("name", ["diamond3d_12x10x16", "fork_join_20x16x24"])
def test_undersized_join_is_reported(name):
    """A join whose branches the consumer reads at different distances (``l[i,j,k] + r[i+1,j,k]``):
    ``compute_delay_buffer`` -- here as in the reference (kernel_chain_graph.py:476-559) -- sizes the
    edge for the difference in path latency only, the near branch backs up and the network stops.
    The reference's simulator reports the same programs as ``buffer ... overflow occurred``."""
    with pytest.raises(SimulationDeadlock) as info:
        _simulate(program_path(name))
    assert "words held" in str(info.value)


def test_interface_mirrors_the_reference():
    # stencilflow/simulator.py:35-39 and the methods run_program / kernel_chain_graph call
    params = list(inspect.signature(Simulator.__init__).parameters)[1:]
    assert params == ["program_name", "program_description", "input_nodes", "kernel_nodes", "output_nodes",
                      "dimensions", "write_output", "log_level"]
    for method in ("simulate", "initialize", "step_execution", "finalize", "get_result", "all_done", "diagnostics"):
        assert callable(getattr(Simulator, method))


def test_program_counters_and_report(capsys):
    path = program_path("ref_simulator9")
    chain = KernelChainGraph(path)
    sim = Simulator("ref_simulator9", helper.parse_json(path), chain.input_nodes, chain.kernel_nodes,
                    chain.output_nodes, chain.dimensions, False, 0)
    sim.initialize()
    steps = 0
    while not sim.all_done():
        assert sim.step_execution() or steps < 3
        steps += 1
    assert steps == sim.cycles
    assert all(k.program_counter == sim.total for k in chain.kernel_nodes.values())
    assert all(i.program_counter == sim.total for i in chain.input_nodes.values())
    assert all(o.program_counter == sim.total for o in chain.output_nodes.values())
    text = sim.report()
    assert "channel" in text and "latency" in text


def test_kernel_chain_graph_command_line(capsys):
    """reference kernel_chain_graph.py:777-817: -stencil_file ... -simulate -report"""
    from stencilflow_b200 import kernel_chain_graph
    chain, sim = kernel_chain_graph.main(["-stencil_file", program_path("ref_simulator12"), "-simulate", "-report",
                                          "-log-level", "0"])
    out = capsys.readouterr().out
    assert sim.all_done() and "total buffer size" in out and "channel kernelA_kernelB" in out
    assert np.allclose(sim.get_result()["res"][:3], [20.25, 20.25, 19.25])
