#!/usr/bin/env python3
"""Where fusion starts to pay: an 8-operator Jacobi-3D chain on N^3 grids, N = 32 ... 256 -- device time of one
execution (CUDA events around 200 back-to-back executions, CUDA graph off) for the planner's choice, fused
passes of 4 operators and eight one-operator launches.  Prints one line per size."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from stencilflow_b200 import build, programs
    build.build_native()
    from stencilflow_b200.cuda_program import CudaProgram
    from stencilflow_b200.planner import PlanOptions
    os.environ["SFB200_GRAPH"] = "0"
    for n in (32, 48, 64, 96, 128, 192, 256):
        prog = programs.jacobi3d_chain([n, n, n], 8)
        path = programs.write_program(prog, "t_%d" % n)
        row = []
        for label, opts in (("planned", None), ("fused4", PlanOptions(max_depth=4)), ("fused2", PlanOptions(max_depth=2)),
                            ("unfused", PlanOptions(fuse=False))):
            p = CudaProgram(path, plan_options=opts, device=0)
            rt = p.rt
            rt.fill_hash(p.buffers["a"].dptr, n ** 3, np.float32, seed=5)
            for _ in range(20):
                p.execute()
            rt.stream_synchronize()
            e0, e1 = rt.event_create(), rt.event_create()
            rt.event_record(e0)
            for _ in range(200):
                p.execute()
            rt.event_record(e1)
            rt.event_synchronize(e1)
            us = 1e3 * rt.elapsed_ms(e0, e1) / 200
            row.append("{} {:7.2f} us ({})".format(label, us, "+".join(str(len(l.ops)) + l.family[0] for l in p.lowered.launches)))
            p.close()
        print("{:>4}^3: ".format(n) + "   ".join(row), flush=True)


if __name__ == "__main__":
    main()
