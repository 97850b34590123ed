#!/usr/bin/env python3
"""Golden OUTPUT vectors produced by the reference's own code, run in the build container.

The reference's compiled paths (DaCe CPU program / FPGA emulation) cannot run here: DaCe 0.10.8 needs
Python < 3.10 and several absent packages.  Its third evaluator, the cycle-level dataflow simulator
(``stencilflow/simulator.py`` driving ``Kernel.try_read/try_execute/try_write``, ``stencilflow/kernel.py:634-738``,
arithmetic by ``stencilflow/calculator.py``), is plain Python and imports DaCe only for the *names* of
the data types.  This script imports those reference modules UNMODIFIED from /root/reference and runs
``KernelChainGraph`` + ``Simulator`` exactly as ``stencilflow/run_program.py:48-61`` does, with two
shims that live in this process only:

* a stub ``dace`` package exposing ``dace.dtypes.{typeclass,float32,float64,...}`` (a numpy-scalar
  type with ``.type``/``.bytes``/``.ctype``/``.to_string()`` -- what ``helper.str_to_dtype``,
  ``BaseKernelNodeClass`` and ``Kernel`` touch);
* ``ast.parse`` re-wraps ``Subscript.slice`` in an ``Index``-like node with a ``.value`` field, the
  pre-3.9 shape the reference reads (``compute_graph_nodes.py:197-218``).

The simulator only handles what the reference's own simulator programs use: 3-D programs whose inputs
are full-dimensional arrays given as lists or files, ``constant`` boundaries, no ``and``/``or``.  (Its
2-D programs stop with a TypeError in ``helper.list_add_cwise`` -- the ``None`` index convention for
absent dimensions post-dates the simulator -- which is why ``run_simulation`` is switched off in
``test/test_stencilflow.py:203-204``.)  Within that envelope CASES covers the reference's own 3-D
test programs and this repository's 3-D constant-boundary test programs (forks/joins, box taps,
ternaries, asymmetric offsets, several inputs) on seeded random inputs.

Nothing of the reference is copied into the repository; only the simulator's outputs are stored
(tests/golden/reference_sim.npz + reference_sim.json).  tests/test_oracle.py requires
oracle/reference_numpy.py to reproduce them and tests/test_parity_gpu.py requires the CUDA path to,
which pins both to outputs of the reference itself.

Run from the repository root (build container only):  python tests/golden/make_reference_sim_golden.py
"""
import ast
import contextlib
import io
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
OUT_NPZ = os.path.join(HERE, "reference_sim.npz")
OUT_JSON = os.path.join(HERE, "reference_sim.json")

# (case name, program file, seed): seed None = the inputs the program file itself names
# ("ref_*" files under tests/programs are the reference's test/stencils/*.json)
CASES = [
    ("ref_simulator12", "ref_simulator12", None),
    ("ref_simulator12_rand", "ref_simulator12", 11),
    ("ref_jacobi3d_32x32x32", "ref_jacobi3d_32x32x32", None),
    ("ref_jacobi3d_32x32x32_8itr_8vec", "ref_jacobi3d_32x32x32_8itr_8vec", None),
    ("ref_jacobi3d_32x32x32_8itr_8vec_rand", "ref_jacobi3d_32x32x32_8itr_8vec", 12),
    ("jacobi3d_16x24x32_5itr_const1", "jacobi3d_16x24x32_5itr_const1", 13),
    ("box3d_10x12x16", "box3d_10x12x16", 14),
    ("synth_box_12x10x16_3st", "synth_box_12x10x16_3st", 16),
    ("synth_fork_16x12x16_5st", "synth_fork_16x12x16_5st", 17),
    ("trig3d_8x10x12_f64", "trig3d_8x10x12_f64", 18),
]
# NOT usable as golden although the simulator runs them: operators with several statements
# (``d = ...; flx = ... d ...``: hdiff_const_10x12x8_f64, multistmt3d_6x8x10_f64).  ``Calculator.evaluate``
# drops everything up to the first "=" and evaluates ``tree.body[0]`` only (calculator.py:164-175), i.e. it
# returns the value of the FIRST statement, which is not what the reference's compiled paths compute
# (the tasklet runs all statements, stencil/cpu.py:141-179).
# Outside the envelope (tried, reported by main() as NOT SIMULATED): fork_join_20x16x24 and
# diamond3d_12x10x16 -- joins of paths of unequal length overflow a delay buffer inside the reference's
# simulator ("RuntimeError: buffer b overflow occurred", bounded_queue.py:122), every 2-D program
# (TypeError in helper.list_add_cwise), programs with 0-D or lower-dimensional inputs
# (bounded_queue.py:68 len() of unsized object), and/or (compute_graph_nodes BoolOp not implemented).


# Cases the reference's simulator cannot run as they are -- ``shrink`` boundaries (kernel.py:534-540 raises
# NotImplementedError), 2-D programs, operators with several statements -- are handed to it in an
# EQUIVALENT form, equivalent by the reference's own definitions, and its output is mapped back:
#   "shrink_const": ``shrink`` is the constant -100000 (stencil/_common.py:8 JUNK_VAL, stencil/cpu.py:91-95);
#   "inline":       an operator ``t = e1; op = e2(t)`` is the operator ``op = e2((e1))``: the tasklet runs its
#                   statements in order and they are pure expressions (stencil/cpu.py:141-179).  (Splitting
#                   it into two operators instead overflows a delay buffer inside the reference's simulator.)
#   "embed2d":      a 2-D program over [Nj, Nk] is the 3-D program over [2, Nj, Nk] without i-offsets (the
#                   reference itself lifts 2-D programs to 3-D, kernel_chain_graph.py:399-403); plane 0 is the result.
# The stored arrays have the ORIGINAL program's shape; tests run the original program and compare the
# interior (``halo`` cells from every border excluded, as ``run_program -halo`` does for shrink programs).
# (case, program, seed, transforms, halo, {input: (lo, hi)})
TRANSFORMED_CASES = [
    ("jacobi3d_shrink_via_const", "jacobi3d_24x20x40_4itr_shrink_f64", 21, ["shrink_const"], 4, {}),
    ("hdiff_shrink_via_const_inline", "hdiff_24x28x16", 22, ["shrink_const", "inline"], 2,
     {"inp": (1.0, 2.0), "coeff": (0.0, 0.05)}),
    ("hdiff_f64_shrink_via_const_inline", "hdiff_16x20x8_f64", 23, ["shrink_const", "inline"], 2,
     {"inp": (1.0, 2.0), "coeff": (0.0, 0.05)}),
    ("multistmt3d_via_inline", "multistmt3d_6x8x10_f64", 24, ["inline"], 0, {}),
    ("jacobi2d_shrink_via_const_embed", "jacobi2d_96x128_6itr_shrink_f64", 25, ["shrink_const", "embed2d"], 6, {}),
    ("ref_jacobi2d_128x128_via_embed", "ref_jacobi2d_128x128", 26, ["embed2d"], 0, {}),
    ("jacobi2d_const_f32_via_embed", "jacobi2d_64x64_4itr_const_f32", 27, ["embed2d"], 0, {}),
]


def transform_program(prog, inputs, transforms):
    """(program', inputs', map_back) for the equivalent forms above.  Uses the real ``ast.parse``."""
    parse = getattr(ast, "_sf_real_parse", ast.parse)
    prog = json.loads(json.dumps(prog))
    inputs = dict(inputs)
    map_back = lambda arr: arr                                          # noqa: E731
    if "shrink_const" in transforms:
        for op in prog["program"].values():
            for bc in op["boundary_conditions"].values():
                if bc["type"] == "shrink":
                    bc["type"], bc["value"] = "constant", -100000.0
    if "inline" in transforms:
        for name, op in prog["program"].items():
            stmts = [st.strip() for st in op["computation_string"].split(";") if st.strip()]
            values = {}
            for stmt in stmts:
                tree = parse(stmt)
                target = tree.body[0].targets[0].id

                class Sub(ast.NodeTransformer):
                    def visit_Name(self, node):
                        return parse("(" + values[node.id] + ")", mode="eval").body if node.id in values else node

                values[target] = ast.unparse(Sub().visit(tree.body[0].value))
            op["computation_string"] = "{} = {}".format(name, values[name])
    if "embed2d" in transforms:
        assert len(prog["dimensions"]) == 2
        prog["dimensions"] = [2] + list(prog["dimensions"])

        class Lift(ast.NodeTransformer):
            def visit_Subscript(self, node):
                elts = node.slice.elts if isinstance(node.slice, ast.Tuple) else [node.slice]
                node.slice = ast.Tuple(elts=[ast.Name(id="i", ctx=ast.Load())] + list(elts), ctx=ast.Load())
                return node

        for op in prog["program"].values():
            stmts = []
            for stmt in op["computation_string"].split(";"):
                if stmt.strip():
                    stmts.append(ast.unparse(Lift().visit(parse(stmt.strip()))))
            op["computation_string"] = "; ".join(stmts)
        for cfg in prog["inputs"].values():
            assert "input_dims" not in cfg
        inputs = {k: np.stack([v, v]) for k, v in inputs.items()}
        map_back = lambda arr: np.asarray(arr).reshape(prog["dimensions"])[0]     # noqa: E731
    return prog, inputs, map_back


def case_program(program):
    """The program description of a case (a dict, as parsed from tests/programs/<program>.json)."""
    with open(os.path.join(ROOT, "tests", "programs", program + ".json")) as f:
        return json.load(f)


def case_inputs(prog, seed, ranges=None):
    """Inputs of a case as {name: ndarray of the program's shape}: U[0.5, 1.5) (or ``ranges[name]``) from
    ``seed``, or (seed None) what the program file names -- ``constant:v`` or a list; .dat files are zeros here."""
    shape = tuple(prog["dimensions"])
    out = {}
    rng = np.random.default_rng(seed) if seed is not None else None
    for name in sorted(prog["inputs"]):
        spec = prog["inputs"][name]
        dt = np.dtype(spec["data_type"]).type
        if rng is not None:
            lo, hi = (ranges or {}).get(name, (0.5, 1.5))
            out[name] = rng.uniform(lo, hi, size=shape).astype(dt)
        elif isinstance(spec["data"], list):
            out[name] = np.array(spec["data"], dtype=dt).reshape(shape)
        elif str(spec["data"]).startswith("constant:"):
            out[name] = np.full(shape, float(spec["data"].split(":")[1]), dtype=dt)
        elif "zeros" in str(spec["data"]):
            out[name] = np.zeros(shape, dtype=dt)
        else:
            raise ValueError("unsupported input specification " + str(spec["data"]))
    return out


def install_shims():
    # ---- stub dace.dtypes -------------------------------------------------------------------
    class typeclass:
        def __init__(self, nptype, cname):
            self.type = nptype
            self.bytes = np.dtype(nptype).itemsize
            self.ctype = cname
            self.dtype = self
            self.veclen = 1

        def __call__(self, *args, **kwargs):
            return self.type(*args, **kwargs)

        def to_string(self):
            return self.type.__name__

        def as_numpy_dtype(self):
            return np.dtype(self.type)

        def __repr__(self):
            return self.type.__name__

    dace = types.ModuleType("dace")
    dtypes = types.ModuleType("dace.dtypes")
    dtypes.typeclass = typeclass
    for name, cname in (("float32", "float"), ("float64", "double"), ("int32", "int"), ("int64", "long long"),
                        ("int8", "char"), ("int16", "short"), ("uint8", "unsigned char"), ("uint32", "unsigned int"),
                        ("uint64", "unsigned long long"), ("bool", "bool")):
        nptype = np.bool_ if name == "bool" else getattr(np, name)
        obj = typeclass(nptype, cname)
        setattr(dtypes, name, obj)
        setattr(dace, name, obj)
    dace.dtypes = dtypes
    sys.modules["dace"] = dace
    sys.modules["dace.dtypes"] = dtypes

    # ---- pre-3.9 shape of Subscript.slice -----------------------------------------------------
    class Index(ast.AST):
        _fields = ("value",)

    real_parse = ast.parse
    ast._sf_real_parse = real_parse

    def parse(source, *args, **kwargs):
        tree = real_parse(source, *args, **kwargs)
        for node in ast.walk(tree):
            if isinstance(node, ast.Subscript) and not isinstance(node.slice, Index):
                node.slice = Index(value=node.slice)
        return tree

    ast.parse = parse

    # ---- the reference package without its __init__ (which imports the DaCe code generators) ----
    pkg = types.ModuleType("stencilflow")
    pkg.__path__ = [os.path.join(REF, "stencilflow")]
    sys.modules["stencilflow"] = pkg
    import importlib
    helper = importlib.import_module("stencilflow.helper")
    for k, v in vars(helper).items():          # `from .helper import *` of the real __init__
        if not k.startswith("_"):
            setattr(pkg, k, v)
    return pkg


def run(prog, inputs, workdir, name):
    """Runs the reference's KernelChainGraph + Simulator on ``prog`` with ``inputs`` written into the
    program file as flat lists (the form the reference's own simulator programs use)."""
    import importlib
    kcg = importlib.import_module("stencilflow.kernel_chain_graph")
    sim_mod = importlib.import_module("stencilflow.simulator")
    helper = importlib.import_module("stencilflow.helper")
    log = importlib.import_module("stencilflow.log_level")
    prog = json.loads(json.dumps(prog))
    for k, v in inputs.items():
        prog["inputs"][k]["data"] = [float(x) for x in np.ravel(v)]
    path = os.path.join(workdir, name + ".json")
    with open(path, "w") as f:
        json.dump(prog, f)
    desc = helper.parse_json(path)
    sink = io.StringIO()
    cells = int(np.prod(prog["dimensions"]))
    with contextlib.redirect_stdout(sink):
        chain = kcg.KernelChainGraph(path=path, plot_graph=False, log_level=log.LogLevel.NO_LOG)
        sim = sim_mod.Simulator(program_name=name, program_description=desc, input_nodes=chain.input_nodes,
                                kernel_nodes=chain.kernel_nodes, output_nodes=chain.output_nodes,
                                dimensions=chain.dimensions, write_output=False, log_level=log.LogLevel.NO_LOG)
        # Simulator.simulate() (simulator.py:186-215) with a cycle cap: a model deadlock must not hang
        sim.initialize()
        cycles = 0
        while not sim.all_done():
            sim.step_execution()
            cycles += 1
            if cycles > 8 * cells + 100000 or "Diagnosis output" in sink.getvalue():
                text = sink.getvalue()
                at = text.find("Traceback")
                raise RuntimeError("the reference simulator stopped making progress:\n" +
                                   (text[at:at + 3000] if at >= 0 else text[-2000:]))
        sim.finalize()
        res = sim.get_result()
    return {k: np.asarray(v) for k, v in res.items()}, cycles


def main():
    """All CASES, or only those named on the command line (merged into the existing files)."""
    import tempfile
    install_shims()
    arrays, index = {}, {}
    only = set(sys.argv[1:])
    if only and os.path.isfile(OUT_NPZ):
        with np.load(OUT_NPZ) as z:
            arrays = {k: z[k] for k in z.files}
        with open(OUT_JSON) as f:
            index = json.load(f)
    with tempfile.TemporaryDirectory() as work:
        for case, program, seed in CASES:
            if only and case not in only:
                continue
            prog = case_program(program)
            inputs = case_inputs(prog, seed)
            try:
                res, cycles = run(prog, inputs, work, case)
            except Exception as exc:  # noqa: BLE001 -- a program outside the simulator's envelope
                lines = [l for l in str(exc).splitlines() if "Error" in l or "Exception" in l]
                print("{:<40} NOT SIMULATED by the reference: {}".format(case, (lines or [str(exc)[:120]])[-1]), flush=True)
                continue
            index[case] = {"program": program, "seed": seed, "cycles": cycles, "outputs": sorted(res)}
            for field, arr in res.items():
                dt = np.dtype(prog["program"][field]["data_type"])
                arrays[case + "/" + field] = np.asarray(arr, dtype=dt).reshape(prog["dimensions"])
            print("{:<40} {:>8} cycles  {}".format(case, cycles, {
                k: float(np.sum(np.asarray(v, dtype=np.float64))) for k, v in res.items()}), flush=True)
        for case, program, seed, transforms, halo, ranges in TRANSFORMED_CASES:
            if only and case not in only:
                continue
            prog = case_program(program)
            inputs = case_inputs(prog, seed, ranges)
            tprog, tinputs, map_back = transform_program(prog, inputs, transforms)
            try:
                res, cycles = run(tprog, tinputs, work, case)
            except Exception as exc:  # noqa: BLE001
                lines = [l for l in str(exc).splitlines() if "Error" in l or "Exception" in l]
                print("{:<40} NOT SIMULATED by the reference: {}".format(case, (lines or [str(exc)[:200]])[-1]), flush=True)
                continue
            outputs = sorted(o for o in res if o in prog["outputs"])
            index[case] = {"program": program, "seed": seed, "cycles": cycles, "outputs": outputs,
                           "transforms": transforms, "halo": halo, "ranges": ranges}
            for field in outputs:
                dt = np.dtype(prog["program"][field]["data_type"])
                arrays[case + "/" + field] = np.asarray(map_back(np.asarray(res[field])), dtype=dt).reshape(prog["dimensions"])
            print("{:<40} {:>8} cycles  {}".format(case, cycles, {
                k: float(np.sum(np.asarray(arrays[case + "/" + k], dtype=np.float64))) for k in outputs}), flush=True)
    np.savez_compressed(OUT_NPZ, **arrays)
    with open(OUT_JSON, "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)
    print("wrote", OUT_NPZ, OUT_JSON)


if __name__ == "__main__":
    main()
