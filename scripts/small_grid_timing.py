#!/usr/bin/env python3
"""Launch-bound case: BASELINE configs[0] (8 chained Jacobi-3D operators on 32^3) -- device time of one
execution for the planner's choice (one-operator kernels in a CUDA graph), the same without the graph,
and fused passes, each checked against the oracle.  Prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from stencilflow_b200 import build, programs, synthetic
    build.build_native()
    from stencilflow_b200.cuda_program import CudaProgram
    from stencilflow_b200.planner import PlanOptions
    from oracle import reference_numpy as rn
    name, prog, _ = programs.baseline_config(0)
    path = programs.write_program(prog, name)
    a = synthetic.fill_hash((32, 32, 32), np.float32, 1234)
    expected = rn.run_reference(prog, {"a": a})["b7"]
    out = {}
    for label, opts, graph in (("planned", None, "1"), ("planned_no_graph", None, "0"),
                               ("fused_depth4", PlanOptions(max_depth=4), "0"),
                               ("fused_depth2", PlanOptions(max_depth=2), "0"),
                               ("unfused_no_graph", PlanOptions(fuse=False), "0")):
        os.environ["SFB200_GRAPH"] = graph
        p = CudaProgram(path, plan_options=opts, device=0)
        res = np.zeros((32, 32, 32), np.float32)
        p(a_host=a, b7_host=res)
        err = float(rn.max_relative_error(expected, res))
        times = p.time_execution(repetitions=200, warmup=20)
        out[label] = {"launches": len(p.lowered.launches), "families": sorted({l.family for l in p.lowered.launches}),
                      "us_median": round(1e3 * float(np.median(times)), 2), "max_rel_err": err}
        p.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
