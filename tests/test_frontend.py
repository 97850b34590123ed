"""Front-end unit tests.  The first three classes restate the reference's own unit tests
(reference test/test_stencilflow.py:17-84 BoundedQueueTest, :114-162 HelperTest) against this package;
the rest pin the analysis quantities the reference computes for its test programs."""
import os

import numpy as np
import pytest

import stencilflow_b200 as sf
from stencilflow_b200 import helper
from stencilflow_b200.bounded_queue import BoundedQueue
from stencilflow_b200.stencil_op import make_program

from conftest import program_path, all_programs


class TestBoundedQueue:
    def test_import(self):
        queue = BoundedQueue(name="test", maxsize=5)
        collection = [1.0, 2.0, 3.0, 4.0, 5.0]
        queue.import_data(collection)
        assert queue.size() == len(collection)
        assert queue.try_peek_last() == collection[-1]
        with pytest.raises(RuntimeError):
            queue.import_data(6 * [1.0])

    def test_enq_deq(self):
        queue = BoundedQueue(name="test", maxsize=1, collection=[1.0])
        assert queue.size() == 1
        assert queue.dequeue() == 1.0
        assert queue.size() == 0 and queue.is_empty()
        with pytest.raises(RuntimeError):
            queue.dequeue()
        queue.enqueue(1.0)
        assert queue.is_full()
        with pytest.raises(RuntimeError):
            queue.enqueue(2.0)

    def test_try_enq_deq(self):
        queue = BoundedQueue(name="test", maxsize=1, collection=[1.0])
        assert queue.try_dequeue() == 1.0
        assert queue.is_empty()
        assert queue.try_dequeue() is False
        assert queue.try_enqueue(1.0) is True
        assert queue.is_full()
        assert queue.try_enqueue(1.0) is False

    def test_peek(self):
        queue = BoundedQueue(name="test", maxsize=2, collection=[1.0, 2.0])
        assert queue.peek(0) == 1.0 and queue.peek(1) == 2.0
        assert queue.try_peek_last() == 2.0
        queue.dequeue()
        queue.dequeue()
        assert queue.try_peek_last() is False

    def test_minimum_capacity(self):
        assert BoundedQueue("q", 0).maxsize == 1


class TestHelper:
    def test_index_math(self):
        assert helper.max_dict_entry_key({"a": [1, 0, 0], "b": [0, 1, 0], "c": [0, 0, 1]}) == "a"
        assert helper.list_add_cwise([1, 2, 3], [3, 2, 1]) == [4, 4, 4]
        assert helper.list_subtract_cwise([1, 2, 3], [1, 2, 3]) == [0, 0, 0]
        assert helper.list_subtract_cwise([None, 2, 3], [None, 1, 1]) == [None, 1, 2]
        assert helper.dim_to_abs_val([3, 2, 1], [10, 10, 10]) == 321
        assert helper.convert_3d_to_1d(dimensions=[10, 10, 10], index=[3, 2, 1]) == 321
        assert helper.convert_3d_to_1d(dimensions=[10, 10, 10], index=[None, 2, 1]) == 21
        assert helper.convert_3d_to_1d(dimensions=[10, 10, 10], index=[]) == 0

    def test_load_array_files(self, tmp_path):
        # reference fixtures helper_test.csv / helper_test.dat hold [7.0, 7.0]
        csv = tmp_path / "helper_test.csv"
        csv.write_text("7.0,7.0\n")
        dat = tmp_path / "helper_test.dat"
        np.array([7.0, 7.0]).tofile(str(dat))
        f64 = helper.str_to_dtype("float64")
        assert list(helper.load_array({"data": str(csv), "data_type": f64})) == [7.0, 7.0]
        assert list(helper.load_array({"data": str(dat), "data_type": f64})) == [7.0, 7.0]
        assert os.path.getsize(str(dat)) == 16

    def test_save_load_roundtrip(self, tmp_path):
        out_data = np.array([1.0, 2.0, 3.0])
        cfg = {"data": str(tmp_path / "test.dat"), "data_type": helper.str_to_dtype("float64")}
        helper.save_array(out_data, cfg["data"])
        assert helper.arrays_are_equal(out_data, helper.load_array(cfg))

    def test_unique(self):
        assert sorted(helper.unique([1.0, 2.0, 1.0])) == [1.0, 2.0]

    def test_generated_inputs(self):
        f32 = helper.str_to_dtype("float32")
        arr = helper.load_array({"data": "constant:0.5", "data_type": f32}, shape=[2, 3])
        assert arr.shape == (2, 3) and arr.dtype == np.float32 and np.all(arr == 0.5)
        assert helper.load_array({"data": "constant:2", "data_type": f32, "input_dims": []}) == 2.0
        with pytest.raises(ValueError):
            helper.load_array({"data": "constant:1", "data_type": f32})
        with pytest.raises(AttributeError):
            helper.str_to_dtype("float1")
        with pytest.raises(TypeError):
            helper.str_to_dtype(7)

    def test_arrays_are_equal(self):
        a = np.array([1.0, -2.0, 0.0], dtype=np.float32)
        assert helper.arrays_are_equal(a, a.copy())
        assert helper.arrays_are_equal(a, a * np.float32(1 + 5e-6))
        assert not helper.arrays_are_equal(a, a * np.float32(1 + 5e-5))
        # tighter than the reference, whose signed divisor lets negative pairs pass (helper.py:273-276)
        assert not helper.arrays_are_equal(np.array([-1.0]), np.array([-2.0]))
        assert not helper.arrays_are_equal(np.array([1.0]), np.array([np.nan]))
        assert helper.arrays_are_equal(np.array([1.0]), np.array([1.0 + 1e-13]), tolerance=1e-12)

    def test_aligned(self):
        raw = np.arange(101, dtype=np.float32)[1:]
        al = helper.aligned(raw, 64)
        assert al.ctypes.data % 64 == 0 and np.array_equal(al, raw)


class TestKernelChainGraph:
    def test_jacobi3d_chain_analysis(self):
        chain = sf.KernelChainGraph(program_path("ref_jacobi3d_32x32x32_8itr_8vec"))
        assert chain.dimensions == [32, 32, 32] and chain.kernel_dimensions == 3
        assert chain.vectorization == 4           # the "_8vec" file really says 4
        assert list(chain.kernel_nodes) == ["b%d" % i for i in range(8)]
        k = chain.kernel_nodes["b3"]
        assert sorted(k.graph.accesses["b2"]) == sorted(
            [[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, -1], [0, -1, 0], [-1, 0, 0]])
        # sliding window of two planes, split at the accesses (reference kernel.py:388-427)
        assert [q.maxsize for q in k.internal_buffer["b2"]] == [992, 31, 2, 31, 992]
        # 1 + mult(16) + 5 adds(80) = 97 cycles, divided by W=4
        assert k.graph.max_latency == 25
        assert k.graph.buffer_size["b2"] == [2, 0, 0 + 3]   # |max_index - min_index| (+W-1 innermost)
        assert k.graph.min_index["b2"] == [-1, 0, 0] and k.graph.max_index["b2"] == [1, 0, 0]
        assert chain.minimum_communication_volume() == 2 * 32 ** 3 * 4
        assert chain.operation_count() == {"Mult": (8, 8 * 32 ** 3), "Add": (40, 40 * 32 ** 3)}
        order = [n.name for n in chain.topological_order()]
        assert order.index("a") < order.index("b0") < order.index("b7")
        assert set(chain.channels) >= {"a_b0", "b0_b1", "b7_b7"}

    def test_jacobi2d_padding(self):
        chain = sf.KernelChainGraph(program_path("ref_jacobi2d_128x128"))
        assert chain.dimensions == [1, 128, 128] and chain.kernel_dimensions == 2
        k = chain.kernel_nodes["b"]
        assert sorted(k.graph.accesses["a"], key=str) == sorted(
            [[None, 1, 0], [None, 0, 1], [None, 0, -1], [None, -1, 0]], key=str)
        assert [q.maxsize for q in k.internal_buffer["a"]] == [127, 2, 127]
        assert k.graph.max_latency == 1 + 16 + 3 * 16

    def test_varying_dimensionality(self):
        chain = sf.KernelChainGraph(program_path("ref_varying_dimensionality"))
        acc = chain.kernel_nodes["out"].graph.accesses
        assert sorted(acc["in1d"], key=str) == sorted([[None, None, 42], [None, None, 0]], key=str)
        assert sorted(acc["in2d"], key=str) == sorted([[1, None, 0], [0, None, 0]], key=str)
        assert acc["in0d"] == [[0, 0, 0]]
        # inputs count at their own dimensionality (kernel_chain_graph.py:749-768)
        assert chain.minimum_communication_volume() == 8 + 32 * 4 + 8 * 32 * 4 + 8 * 16 * 32 * 8 + 8 * 16 * 32 * 4

    def test_fork_join_delay_buffers(self):
        chain = sf.KernelChainGraph(program_path("ref_simulator9"))
        res = chain.kernel_nodes["res"]
        # kernelA reaches res directly and through kernelB: the direct edge is delayed by B's latency
        assert res.delay_buffer["kernelB"].maxsize == 1
        assert res.delay_buffer["kernelA"].maxsize == chain.kernel_nodes["kernelB"].graph.max_latency + 2

    def test_cycle_detected(self, tmp_path):
        prog = {"inputs": {"a": {"data": "constant:1", "data_type": "float32"}}, "outputs": ["x"],
                "dimensions": [4, 4, 4],
                "program": {
                    "x": {"computation_string": "x = y[i,j,k] + a[i,j,k]", "boundary_conditions": {},
                          "data_type": "float32"},
                    "y": {"computation_string": "y = x[i,j,k]", "boundary_conditions": {},
                          "data_type": "float32"}}}
        import json
        p = tmp_path / "cycle.json"
        p.write_text(json.dumps(prog))
        with pytest.raises(ValueError, match="Cycle detected"):
            sf.KernelChainGraph(str(p))

    @pytest.mark.parametrize("name", all_programs())
    def test_every_program_builds_operators(self, name):
        chain = sf.KernelChainGraph(program_path(name))
        prog = make_program(chain)
        assert [op.name for op in prog.ops][-1] in chain.program
        assert set(prog.outputs) == set(chain.outputs)
        for op in prog.ops:
            for field, (mask, offs) in op.accesses.items():
                assert len(mask) == len(prog.iterators)
                assert all(len(o) == sum(mask) for o in offs)

    def test_vectorization_must_divide(self, tmp_path):
        import json
        with open(program_path("ref_jacobi2d_128x128")) as f:
            prog = json.load(f)
        prog["dimensions"] = [128, 130]
        prog["vectorization"] = 4
        p = tmp_path / "bad.json"
        p.write_text(json.dumps(prog))
        with pytest.raises(ValueError, match="not divisible"):
            make_program(sf.KernelChainGraph(str(p)))


def test_exported_program_conventions_parse():
    """JSON as ``sdfg_to_stencilflow`` emits it (reference sdfg_to_stencilflow.py:522-767): versioned names,
    ``constants`` with string values and a data type, ``btype`` keys, astunparse's newline-separated
    statements, ``input_dims`` on every input, the J,K,I layout (vertical axis in the middle)."""
    from stencilflow_b200.kernel_chain_graph import KernelChainGraph
    from stencilflow_b200.stencil_op import make_program
    chain = KernelChainGraph(program_path("sdfgexport_hdiff_jki_48x8x64_f64"))
    assert chain.dimensions == [48, 8, 64]
    assert set(chain.kernel_nodes) == {"lap", "flx", "fly", "out__1", "out"}
    assert chain.constants["dcoef"]["value"] == "0.25"
    prog = make_program(chain)
    ops = {op.name: op for op in prog.ops}
    # the versioned intermediate feeds the final write of the same field
    assert "out__1" in ops["out"].accesses and prog.fields["out"].kind == "output"
    assert prog.fields["out__1"].kind == "intermediate"
    # two statements, the second using the first's temporary
    assert [s.target for s in ops["flx"].statements] == ["d", "flx"]
    # boundary keys given as btype
    assert ops["lap"].boundary_conditions["inp"]["btype"] == "shrink"
    # the 1-D input lives on the vertical (middle) axis only; the constant is not a field
    assert list(prog.fields["wgt"].dims) == ["j"] and prog.fields["wgt"].shape == (8,)
    assert "dcoef" in ops["out"].scalars and "dcoef" not in prog.fields
    # horizontal offsets sit on the first and the last iterator (J and I of the COSMO layout)
    taps = ops["lap"].offsets3("inp")
    assert sorted(tuple(t) for t in taps) == sorted([(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 0, 1), (0, 0, -1)])
