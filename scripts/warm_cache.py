#!/usr/bin/env python3
"""Compiles (NVRTC, no GPU needed) every plan variant the GPU parity tests force, so that the
content-addressed program cache (.sfcache/, which travels to the GPU box) already holds their cubins
and the box spends its time running kernels instead of compiling them."""
import itertools
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_parity_gpu as T  # noqa: E402
from conftest import all_programs, program_path  # noqa: E402
from stencilflow_b200.cuda_program import CudaProgram  # noqa: E402
from stencilflow_b200.planner import PlanOptions  # noqa: E402


def main():
    t0 = time.time()
    jobs = []
    for name in all_programs():
        jobs.append((program_path(name), PlanOptions(fuse=False)))
        jobs.append((program_path(name), None))
        jobs.append((program_path(name), PlanOptions(max_depth=8)))
    names3 = ["ref_jacobi3d_32x32x32_8itr_8vec", "jacobi3d_16x24x32_5itr_const1", "jacobi3d_24x20x40_4itr_shrink_f64",
              "hdiff_24x28x16", "fork_join_20x16x24", "box3d_10x12x16"]
    for name, v in itertools.product(names3, T.PLAN_VARIANTS):
        jobs.append((program_path(name), PlanOptions(max_depth=v[0], rows_per_thread=v[1], warps=v[2], vector=v[3],
                                                     threads_per_row=v[4] if len(v) > 4 else 0)))
    for name, v in itertools.product([n for n in names3 if n != "hdiff_24x28x16"], T.PAIR_VARIANTS_3D):
        jobs.append((program_path(name), PlanOptions(max_depth=v[0], rows_per_thread=v[1], warps=v[2],
                                                     threads_per_row=v[3], prefetch=v[4], sync=v[5])))
    for name, v in itertools.product(["ref_jacobi3d_32x32x32_8itr_8vec", "fork_join_20x16x24", "box3d_10x12x16",
                                      "jacobi3d_16x24x32_5itr_const1"], T.DIRECT_VARIANTS_3D):
        jobs.append((program_path(name), PlanOptions(max_depth=v[0], rows_per_thread=v[1], warps=v[2],
                                                     threads_per_row=v[3], prefetch=v[4], direct=1)))
    names2 = ["jacobi2d_96x128_6itr_shrink_f64", "jacobi2d_64x64_4itr_const_f32", "ref_jacobi2d_128x128"]
    for name, v in itertools.product(names2, [(1, 8, 0), (2, 8, 0), (4, 8, 0), (6, 16, 0), (4, 8, 4), (2, 16, 8), (3, 8, 8)]):
        jobs.append((program_path(name), PlanOptions(max_depth=v[0], warps=v[1], vector=v[2])))
    done = fresh = 0
    for path, opts in jobs:
        try:
            p = CudaProgram(path, plan_options=opts, allocate=False)
            fresh += 0 if p.was_cached else 1
            done += 1
        except Exception as exc:
            print("skip", os.path.basename(path), str(exc)[:80])
    print("warmed {} variants ({} compiled) in {:.0f} s".format(done, fresh, time.time() - t0))


if __name__ == "__main__":
    main()
