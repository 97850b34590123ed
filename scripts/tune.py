#!/usr/bin/env python3
"""Measures candidate plans of a program on the GPU and records the fastest in
``stencilflow_b200/tuned_plans.json`` (consulted by ``planner.plan_program`` when no knob is set).

    python scripts/tune.py --config 1 [--config 2 ...]      # BASELINE.json configs
    python scripts/tune.py path/to/program.json

A candidate is a (max fusion depth, rows per thread, warps) triple; candidates whose resulting plan
repeats an earlier one are skipped.  Timing: CUDA events around 3 executions after 2 warm-ups, inputs
resident (the same measurement as bench.py's ``value``).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from stencilflow_b200 import build, planner, programs  # noqa: E402
from stencilflow_b200.cuda_program import CudaProgram  # noqa: E402


def candidates(prog):
    ndim = len(prog["dimensions"])
    nops = len(prog["program"])
    depths = sorted({d for d in (1, 2, 3, 4, 6, 8) if d <= nops} | ({nops} if nops <= 8 else set()))
    out = []
    if ndim == 3:
        narrow = prog["dimensions"][-1] < 128
        for d in depths:
            if narrow:
                for r, w in ((2, 0), (3, 0), (4, 0), (2, 8), (2, 12), (3, 8), (4, 8), (1, 0)):
                    out.append((d, r, w))
            else:
                for r, w in ((4, 16), (4, 12), (4, 8), (3, 16), (3, 12), (2, 16)):
                    out.append((d, r, w))
    else:
        depths = sorted(set(depths) | ({16} if nops >= 16 else set()))
        for d in depths:
            for w in (8, 16):
                for v in (0, 4) if d >= 4 else (0,):
                    out.append((d, v, w))          # second slot = cells per thread in 2-D programs
    return out


def measure(path, opts, steps=3, warmup=2):
    program = CudaProgram(path, plan_options=opts, device=0)
    try:
        bench.fill_inputs(program)
        rtm = program.rt
        for _ in range(warmup):
            program.execute()
        rtm.stream_synchronize()
        e0, e1 = rtm.event_create(), rtm.event_create()
        rtm.event_record(e0)
        for _ in range(steps):
            program.execute()
        rtm.event_record(e1)
        rtm.event_synchronize(e1)
        ms = rtm.elapsed_ms(e0, e1) / steps
        sig = [(l.family, len(l.ops), l.info.get("R"), tuple(l.info.get("warps", ())), l.info.get("threads_per_row"))
               for l in program.lowered.launches]
        return ms, sig
    finally:
        program.close()


def tune(name, prog, table, budget_s):
    path = programs.write_program(prog, name)
    key = planner.structure_key(CudaProgram(path, allocate=False).program)
    seen, results = set(), []
    t0 = time.time()
    for (d, r, w) in candidates(prog):
        if time.time() - t0 > budget_s:
            print("  (time budget reached)")
            break
        if len(prog["dimensions"]) == 2:
            opts = planner.PlanOptions(max_depth=d, vector=r, warps=w)
        else:
            opts = planner.PlanOptions(max_depth=d, rows_per_thread=r, warps=w)
        try:
            probe = CudaProgram(path, plan_options=opts, allocate=False)
        except Exception as exc:                       # candidate does not lower
            print("  depth {} rows {} warps {}: {}".format(d, r, w, str(exc)[:80]))
            continue
        sig = tuple((l.family, len(l.ops), l.info.get("R"), l.info.get("V"), tuple(l.info.get("warps", ())),
                     l.info.get("threads_per_row")) for l in probe.lowered.launches)
        if sig in seen:
            continue
        seen.add(sig)
        try:
            ms, _ = measure(path, opts)
        except Exception as exc:
            print("  depth {} rows {} warps {}: failed: {}".format(d, r, w, str(exc)[:120]))
            continue
        results.append((ms, d, r, w, sig))
        print("  depth {} rows {} warps {:>2}: {:8.3f} ms  {}".format(d, r, w, ms, [s[:4] for s in sig][:4]), flush=True)
    if not results:
        return
    ms, d, r, w, sig = min(results)
    nops = len(prog["program"])
    cells = float(np.prod(prog["dimensions"]))
    chosen = ({"max_depth": d, "vector": r, "warps": w} if len(prog["dimensions"]) == 2
              else {"max_depth": d, "rows_per_thread": r, "warps": w})
    entries = [e for e in table.get(key, []) if e.get("shape") != list(prog["dimensions"])]
    table[key] = entries + [{"program": name, "shape": list(prog["dimensions"]), "options": chosen,
                  "measured": {"ms": round(ms, 4), "cell_updates_per_s": nops * cells / (ms * 1e-3),
                               "candidates": len(results), "device": "B200"}}]
    print("  best: depth {} rows {} warps {} -> {:.3f} ms ({:.3e} updates/s)".format(
        d, r, w, ms, nops * cells / (ms * 1e-3)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, action="append", default=[])
    ap.add_argument("--budget", type=float, default=240.0, help="seconds per program")
    ap.add_argument("--out", default=planner.TUNED_PLANS)
    ap.add_argument("programs", nargs="*")
    args = ap.parse_args()
    build.build_native()
    table = planner.load_tuned()
    for c in args.config:
        name, prog, _ = programs.baseline_config(c)
        print("config", c, name, flush=True)
        tune(name, prog, table, args.budget)
    for p in args.programs:
        with open(p) as f:
            prog = json.load(f)
        print(p, flush=True)
        tune(os.path.splitext(os.path.basename(p))[0], prog, table, args.budget)
    with open(args.out, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
