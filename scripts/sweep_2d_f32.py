#!/usr/bin/env python3
"""CTA width for a float32 2-D chain (16 operators, 32768 x 32768, constant boundary): does what was measured on
the float64 chain -- one-warp CTAs with short chunks -- carry over?  Prints ms per step per variant."""
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
    warm = "--warm" in sys.argv
    prog = programs.jacobi2d_chain([32768, 32768], 16, dtype="float32", boundary={"type": "constant", "value": 0.0})
    path = programs.write_program(prog, "jacobi2d_32768_16itr_f32_const")
    n = 32768 * 32768
    first = None
    for label, opts in (("planned", None), ("d8 v8 w1", PlanOptions(max_depth=8, vector=8, warps=1)),
                        ("d8 v8 w2", PlanOptions(max_depth=8, vector=8, warps=2)),
                        ("d8 v8 w4", PlanOptions(max_depth=8, vector=8, warps=4)),
                        ("d8 v4 w1", PlanOptions(max_depth=8, vector=4, warps=1)),
                        ("d8 v4 w2", PlanOptions(max_depth=8, vector=4, warps=2))):
        try:
            p = CudaProgram(path, plan_options=opts, allocate=not warm, device=None if warm else 0)
        except Exception as exc:
            print("{:<10} does not lower: {}".format(label, str(exc)[:80]))
            continue
        l = p.lowered.launches[0]
        desc = "{} launches of {} ops, V={} warps={} tile={}".format(len(p.lowered.launches), len(l.ops), l.info.get("V"),
                                                                      l.info.get("warps"), l.info.get("tile"))
        if warm:
            print(label, desc)
            continue
        rt = p.rt
        rt.fill_hash(p.buffers["a"].dptr, n, np.float32, seed=3)
        for _ in range(2):
            p.execute()
        rt.stream_synchronize()
        e0, e1 = rt.event_create(), rt.event_create()
        rt.event_record(e0)
        for _ in range(5):
            p.execute()
        rt.event_record(e1)
        rt.event_synchronize(e1)
        ms = rt.elapsed_ms(e0, e1) / 5
        s = rt.checksum(p.buffers["b15"].dptr, n, np.float32)[1]
        first = s if first is None else first
        print("{:<10} {:8.3f} ms  {:.3e} upd/s  {}  {}".format(label, ms, 16 * n / (ms * 1e-3), desc,
                                                              "bits==first" if s == first else "BITS DIFFER"), flush=True)
        p.close()


if __name__ == "__main__":
    main()
