"""Multi-GPU driver: the ``cuda`` counterpart of the reference's ``bin/run_distributed_program.py``
(per-rank program, barrier, run, last rank compares with the CPU program, ``:283-341``), with slab
decomposition over the GPUs of one box instead of the operator-pipeline cut of ``split_sdfg``."""

import copy
import os
import re

import numpy as np

from . import helper
from .log_level import LogLevel
from .run_program import _load_reference_backend, tolerance_for


def run_distributed_program(stencil_file, mode="cuda", compare_to_reference=False, input_directory=None,
                            halo=0, repetitions=1, log_level=LogLevel.BASIC, comm=None):
    """Runs on every rank (one process per GPU).  Returns 0 on the verifying rank when the results
    match, raises ``ValueError("Result mismatch.")`` otherwise."""
    from . import distributed
    if mode != "cuda":
        raise ValueError("Unrecognized execution mode: {}".format(mode))
    if isinstance(log_level, int):
        log_level = LogLevel(log_level)
    comm = comm or distributed.make_comm()
    rank, world = comm.rank, comm.world
    verbose = log_level >= LogLevel.BASIC
    description = helper.parse_json(stencil_file)
    name = re.match(r"(.+)\.[^\.]+", os.path.basename(stencil_file)).group(1).replace(".", "_")
    if verbose:
        print("Rank {}/{}: building program {}...".format(rank, world, name), flush=True)
    program = distributed.SlabProgram(stencil_file, comm, device=int(os.environ.get("LOCAL_RANK", rank)))
    if input_directory is None:
        input_directory = os.path.dirname(stencil_file)
    inputs = helper.load_input_arrays(copy.deepcopy(description["inputs"]), prefix=input_directory,
                                      shape=description["dimensions"])
    scalars = {}
    for key, val in inputs.items():
        f = program.program.fields[key]
        if f.is_scalar:
            scalars[key] = val
        else:
            arr = np.asarray(val)
            n = int(np.prod(f.shape))
            program.upload_global(key, np.ascontiguousarray(arr).ravel()[:n].reshape(f.shape))
    if scalars:
        program.set_scalars(scalars)
    comm.barrier()
    for rep in range(max(1, repetitions)):
        if verbose and rank == 0:
            print("Executing CUDA program on {} GPUs (repetition {}/{})...".format(world, rep + 1, repetitions))
        program.execute()
    program.rt.stream_synchronize()
    comm.barrier()
    if verbose:
        print("Rank {} finished (slab {}, {} halo pushes per execution).".format(
            rank, program.slab, len(program.sends)), flush=True)
    outputs = {out: program.gather(out).reshape(description["dimensions"]) for out in program.program.outputs}
    program.close()
    result = None
    if rank == world - 1:
        folder = os.path.join("results", name)
        os.makedirs(folder, exist_ok=True)
        trimmed = {k: (v[tuple(slice(halo, -halo) for _ in v.shape)] if halo > 0 else v) for k, v in outputs.items()}
        helper.save_output_arrays(trimmed, folder)
        if compare_to_reference:
            print("Executing reference program...")
            ref = _load_reference_backend().CompiledReference(stencil_file)
            ref_out = {k: np.zeros_like(v) for k, v in outputs.items()}
            ref(**{k: v for k, v in inputs.items()}, **ref_out)
            print("Comparing to reference program...")
            for k, got in trimmed.items():
                exp = ref_out[k][tuple(slice(halo, -halo) for _ in got.shape)] if halo > 0 else ref_out[k]
                if not helper.arrays_are_equal(np.ravel(exp), np.ravel(got), tolerance=tolerance_for(got.dtype)):
                    print("Expected: {}".format(exp))
                    print("Got:      {}".format(got))
                    raise ValueError("Result mismatch.")
            print("Results verified.")
            result = 0
    comm.barrier()
    return result
