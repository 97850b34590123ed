"""Worker of the world_size-2 CPU test (gloo): executes the slab halo schedule of a program with the
numpy oracle standing in for the kernels and checks the gathered result against the global oracle.
Run by tests/test_distributed_cpu.py; needs RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_numpy as rn  # noqa: E402
from stencilflow_b200 import distributed  # noqa: E402
from stencilflow_b200.cuda_program import CudaProgram  # noqa: E402
from stencilflow_b200.planner import PlanOptions  # noqa: E402


def sub_program(prog, launch, lowered, nplanes):
    """The operators of one launch as a stand-alone program over ``nplanes`` planes of the slab axis."""
    axis = lowered.slab_axis - (3 - len(prog["dimensions"]))
    dims = list(prog["dimensions"])
    dims[axis] = nplanes
    sub = {"dimensions": dims, "outputs": list(launch.writes), "inputs": {}, "program": {}}
    for f in launch.reads:
        cfg = prog["inputs"].get(f) or {"data_type": prog["program"][f]["data_type"]}
        sub["inputs"][f] = {"data": "constant:0", "data_type": cfg["data_type"]}
        if "input_dims" in cfg:
            sub["inputs"][f]["input_dims"] = cfg["input_dims"]
    for name, cfg in prog["inputs"].items():
        if cfg.get("input_dims") == []:
            sub["inputs"][name] = dict(cfg)
    for op in launch.ops:
        sub["program"][op] = prog["program"][op]
    if "constants" in prog:
        sub["constants"] = prog["constants"]
    return sub


def main():
    path, fuse, seed = sys.argv[1], sys.argv[2] == "1", int(sys.argv[3])
    comm = distributed.TorchComm("gloo")
    prog = rn.load_program(path)
    info = rn.ProgramInfo(prog)
    # fusion requested explicitly: the cost model prefers one-operator kernels on grids this small
    plan = CudaProgram(path, allocate=False, plan_options=PlanOptions(fuse=fuse, max_depth=4 if fuse else None))
    lowered = plan.lowered
    axis = lowered.slab_axis
    n = plan.program.shape3[axis]
    halo = distributed.halo_depth(lowered)
    slab = distributed.Slab(comm.rank, comm.world, n, halo)
    sends = distributed.halo_schedule(lowered, slab)
    it = "ijk"[axis]
    rng = np.random.default_rng(seed)           # same seed on every rank: same global inputs
    global_inputs, scalars = {}, {}
    for name in info.inputs:
        shape = info.field_shape(name)
        dt = info.field_type(name)
        if len(shape) == 0:
            scalars[name] = dt(rng.uniform(0.5, 1.5))
        else:
            global_inputs[name] = rng.uniform(0, 1, size=shape).astype(dt)
    expected = rn.run_reference(prog, dict(global_inputs, **scalars))

    def sharded(name):
        return it in info.field_dims(name)

    local = {}
    for name, arr in global_inputs.items():
        local[name] = arr[slab.alloc_begin:slab.alloc_end].copy() if sharded(name) else arr
    nplanes = slab.alloc_end - slab.alloc_begin
    for idx, launch in enumerate(lowered.launches):
        sub = sub_program(prog, launch, lowered, nplanes)
        ins = {f: local[f] for f in launch.reads}
        ins.update(scalars)
        out = rn.run_reference(sub, ins)
        for f in launch.writes:
            local[f] = out[f]
        mine = [(s.peer, s.field, s.src_begin, s.src_end,
                 local[s.field][s.src_begin - slab.alloc_begin:s.src_end - slab.alloc_begin].copy())
                for s in sends if s.launch == idx]
        for src_rank, msgs in enumerate(comm.allgather(mine)):
            for (peer, field, b, e, data) in msgs:
                if peer == comm.rank:
                    local[field][b - slab.alloc_begin:e - slab.alloc_begin] = data
    ok = True
    for name in info.outputs:
        own = local[name][slab.begin - slab.alloc_begin:slab.end - slab.alloc_begin]
        full = np.concatenate(comm.allgather(own), axis=0)
        if not np.array_equal(full, expected[name]):
            ok = False
            print("rank", comm.rank, "MISMATCH", name, float(np.max(np.abs(full - expected[name]))), flush=True)
    summary = {"rank": comm.rank, "ok": ok, "halo": halo, "slab": [slab.begin, slab.end, slab.alloc_begin, slab.alloc_end],
               "sends": len(sends), "launches": len(lowered.launches), "max": comm.max_float(comm.rank + 0.5)}
    print("RESULT " + json.dumps(summary), flush=True)
    comm.barrier()
    comm.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
