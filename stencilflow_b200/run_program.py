"""Driver: load a JSON stencil program, build it, run it on the GPU, compare with the CPU program.

Same signature, prints, result files and return values as the reference driver
(``stencilflow/run_program.py:19-250``) with one more ``mode``: ``"cuda"``.  ``"emulation"`` and
``"hardware"`` are the reference's FPGA modes and are not part of this backend.

Reference comparison (``-compare-to-reference``): the reference generates and runs a CPU SDFG
(``generate_reference``); here the CPU program is the restatement kept under ``oracle/`` (test
infrastructure), used strictly as the checker -- nothing computed on the CPU ever reaches the
``cuda`` result arrays.
"""

import copy
import os
import re

import numpy as np

from . import helper
from .kernel_chain_graph import KernelChainGraph
from .log_level import LogLevel


def _load_reference_backend():
    try:
        from oracle import reference_cpp
        return reference_cpp
    except ImportError:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        if root not in sys.path:
            sys.path.insert(0, root)
        from oracle import reference_cpp
        return reference_cpp


def tolerance_for(dtype) -> float:
    """Maximum relative error accepted by ``-compare-to-reference``: 1e-5 for float32
    (the reference's ``arrays_are_equal`` default, helper.py:261) and 1e-12 for float64."""
    return 1e-12 if np.dtype(dtype) == np.float64 else 1e-5


def run_program(stencil_file,
                mode,
                run_simulation=False,
                compare_to_reference=False,
                input_directory=None,
                use_cached_sdfg=None,
                skip_execution=False,
                generate_input=False,
                synthetic_reads=None,
                specialize_scalars=False,
                plot=False,
                halo=0,
                repetitions=1,
                log_level=LogLevel.BASIC,
                print_result=False,
                xilinx=False):
    if isinstance(log_level, int):
        log_level = LogLevel(log_level)
    if mode != "cuda":
        if mode in ("emulation", "hardware"):
            raise ValueError("Execution mode '{}' is the reference's FPGA path; this backend "
                             "provides mode 'cuda'".format(mode))
        raise ValueError("Unrecognized execution mode: {}".format(mode))
    if xilinx:
        print("Note: -xilinx has no effect in cuda mode.")

    program_description = helper.parse_json(stencil_file)
    name = os.path.basename(stencil_file)
    name = re.match(r"(.+)\.[^\.]+", name).group(1).replace(".", "_")

    if log_level >= LogLevel.BASIC:
        print("Creating kernel graph...")
    chain = KernelChainGraph(path=stencil_file, plot_graph=plot, log_level=log_level)

    simulation_result = None
    if run_simulation:
        # reference run_program.py:48-61
        if log_level >= LogLevel.BASIC:
            print("Running simulation...")
        from .simulator import Simulator
        sim_description = program_description
        if input_directory is not None:
            sim_description = dict(program_description, path=input_directory)
        sim = Simulator(program_name=name, program_description=sim_description,
                        input_nodes=chain.input_nodes, kernel_nodes=chain.kernel_nodes,
                        output_nodes=chain.output_nodes, dimensions=chain.dimensions,
                        write_output=False, log_level=log_level)
        sim.simulate()
        simulation_result = sim.get_result()

    from .cuda_program import CudaProgram

    if log_level >= LogLevel.BASIC:
        print("Generating CUDA program...")
    input_description = copy.deepcopy(program_description["inputs"])
    specialize = None
    if specialize_scalars:
        specialize = {}
        for k, v in input_description.items():
            dims = v.get("input_dims", v.get("dimensions"))
            if dims is not None and len(dims) == 0:
                specialize[k] = helper.load_array(v, prefix=input_directory or os.path.dirname(stencil_file),
                                                  shape=[])
    if log_level >= LogLevel.BASIC:
        print("Compiling CUDA program...")
    execute = not (skip_execution or repetitions == 0)
    program = CudaProgram(chain=chain, log_level=log_level, specialize_scalars=specialize,
                          synthetic_reads=synthetic_reads, allocate=execute)
    if log_level >= LogLevel.BASIC and use_cached_sdfg:
        print("Cached build {}.".format("reused" if program.was_cached else "not found; compiled"))
    reference_program = None
    if compare_to_reference:
        if log_level >= LogLevel.BASIC:
            print("Compiling reference program...")
        reference_program = _load_reference_backend().CompiledReference(stencil_file)

    if not execute:
        if log_level >= LogLevel.BASIC:
            print("Skipping execution and exiting.")
        return

    if log_level >= LogLevel.BASIC:
        print("Loading input arrays...")
    if input_directory is None:
        input_directory = os.path.dirname(stencil_file)
    if generate_input:
        for k in input_description:
            input_description[k]["data"] = "constant:0.5"
    input_arrays = helper.load_input_arrays(input_description, prefix=input_directory,
                                            shape=program_description["dimensions"])

    if log_level >= LogLevel.BASIC:
        print("Initializing output arrays...")
    output_arrays = {
        arr_name: helper.aligned(
            np.zeros(program_description["dimensions"],
                     dtype=program_description["program"][arr_name]["data_type"].type), 64)
        for arr_name in program_description["outputs"]
    }
    if compare_to_reference:
        reference_output_arrays = copy.deepcopy(output_arrays)

    cuda_args = {
        (key + "_host" if hasattr(val, "shape") and len(val.shape) > 0 else key): val
        for key, val in list(input_arrays.items()) + list(output_arrays.items())
    }
    if repetitions == 1:
        print("Executing CUDA program...")
        program(**cuda_args)
        print("Finished running program.")
    else:
        for i in range(repetitions):
            print("Executing repetition {}/{}...".format(i + 1, repetitions))
            program(**cuda_args)
            print("Finished running program.")
    program.close()

    if print_result:
        for key, val in output_arrays.items():
            print(key + ":", val)

    if compare_to_reference:
        print("Executing reference program...")
        ref_inputs = {k: v for k, v in input_arrays.items()}
        if synthetic_reads is not None:
            for k, v in ref_inputs.items():
                if hasattr(v, "shape") and len(v.shape) > 0:
                    ref_inputs[k] = np.full_like(v, synthetic_reads)
        reference_program(**ref_inputs, **reference_output_arrays)
        print("Finished running program.")
        if print_result:
            for key, val in reference_output_arrays.items():
                print(key + ":", val)

    output_folder = os.path.join("results", name)
    os.makedirs(output_folder, exist_ok=True)
    if halo > 0:
        for k, v in output_arrays.items():
            output_arrays[k] = v[tuple(slice(halo, -halo) for _ in v.shape)]
        if compare_to_reference:
            for k, v in reference_output_arrays.items():
                reference_output_arrays[k] = v[tuple(slice(halo, -halo) for _ in v.shape)]
    helper.save_output_arrays(output_arrays, output_folder)
    print("Results saved to " + output_folder)
    if compare_to_reference:
        reference_folder = os.path.join(output_folder, "reference")
        os.makedirs(reference_folder, exist_ok=True)
        helper.save_output_arrays(reference_output_arrays, reference_folder)
        print("Reference results saved to " + reference_folder)

    if compare_to_reference:
        print("Comparing to reference program...")
        for outp in output_arrays:
            got = output_arrays[outp]
            expected = reference_output_arrays[outp]
            if not helper.arrays_are_equal(np.ravel(expected), np.ravel(got),
                                           tolerance=tolerance_for(got.dtype)):
                print("Expected: {}".format(expected))
                print("Got:      {}".format(got))
                raise ValueError("Result mismatch.")
        print("Results verified.")
        if simulation_result is None:
            return 0

    # Compare simulation result to the device result (reference run_program.py:232-250; there it is
    # skipped when -compare-to-reference already returned, here both checks run)
    if simulation_result is not None:
        print("Comparing simulation results...")
        all_match = True
        for outp in output_arrays:
            got = output_arrays[outp]
            simulated = simulation_result[outp].reshape(program_description["dimensions"])
            if halo > 0:
                simulated = simulated[tuple(slice(halo, -halo) for _ in simulated.shape)]
            if print_result:
                print("CUDA result:\n\t{}".format(np.ravel(got)))
                print("Simulation result:\n\t{}".format(np.ravel(simulated)))
            if not helper.arrays_are_equal(np.ravel(simulated), np.ravel(got),
                                           tolerance=tolerance_for(got.dtype)):
                all_match = False
        if all_match:
            print("Results verified.")
            return 0
        print("Result mismatch.")
        return 1
