#!/usr/bin/env python3
"""Times explicit plan variants of a BASELINE config on the GPU and checks that every variant
produces bit-identical output (on-device checksum of the program outputs against the first variant).

    python scripts/sweep_variants.py --config 1 d4r4w8 d4r4w8p5 d4r4w8p5s ...

A variant is written d<depth>[r<rows>][v<cells>][w<warps>][k<threads per row>][p<prefetch>][s][x]
where a trailing ``s`` selects neighbour-only ("pair") synchronisation and ``x`` direct reads of the
input's neighbour rows from the TMA ring.
"""
import argparse
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from stencilflow_b200 import build, planner, programs  # noqa: E402
from stencilflow_b200.cuda_program import CudaProgram  # noqa: E402


def parse(text):
    m = re.fullmatch(r"d(\d+)(?:r(\d+))?(?:v(\d+))?(?:w(\d+))?(?:k(\d+))?(?:p(\d+))?(s?)(x?)", text)
    if not m:
        raise SystemExit("bad variant " + text)
    d, r, v, w, k, p, s, x = m.groups()
    return planner.PlanOptions(max_depth=int(d), rows_per_thread=int(r or 0), vector=int(v or 0),
                               warps=int(w or 0), threads_per_row=int(k or 0), prefetch=int(p or 0),
                               sync="pair" if s else "cta", direct=1 if x else 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("variants", nargs="+")
    args = ap.parse_args()
    build.build_native()
    name, prog, _ = programs.baseline_config(args.config)
    path = programs.write_program(prog, name)
    nops = len(prog["program"])
    cells = float(np.prod(prog["dimensions"]))
    first = None
    for text in args.variants:
        try:
            p = CudaProgram(path, plan_options=parse(text), device=0)
        except Exception as exc:
            print("{:<16} does not lower: {}".format(text, str(exc)[:100]), flush=True)
            continue
        try:
            bench.fill_inputs(p)
            rt = p.rt
            for _ in range(2):
                p.execute()
            rt.stream_synchronize()
            e0, e1 = rt.event_create(), rt.event_create()
            rt.event_record(e0)
            for _ in range(args.steps):
                p.execute()
            rt.event_record(e1)
            rt.event_synchronize(e1)
            ms = rt.elapsed_ms(e0, e1) / args.steps
            sums = []
            for oname, f in p.program.fields.items():
                if f.kind == "output":
                    n = int(np.prod(p.local_shape(oname)))
                    sums.append(p.rt.checksum(p.buffers[oname].dptr, n, f.data_type.type)[1])
            if first is None:
                first = sums
            info = p.lowered.launches[0].info
            regs = ""
            print("{:<16} {:8.3f} ms  {:.3e} upd/s  launches {}  sync {} P {} tile {} smem {}  {}".format(
                text, ms, nops * cells / (ms * 1e-3), [len(l.ops) for l in p.lowered.launches][:6],
                info.get("sync"), info.get("prefetch"), info.get("tile"), p.lowered.launches[0].smem,
                "bits==first" if sums == first else "BITS DIFFER " + str(sums) + " vs " + str(first)), flush=True)
        except Exception as exc:
            print("{:<16} failed: {}".format(text, str(exc)[:200]), flush=True)
        finally:
            p.close()


if __name__ == "__main__":
    main()
