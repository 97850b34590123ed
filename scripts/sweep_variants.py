#!/usr/bin/env python3
"""Times explicit plan variants of a BASELINE config on the GPU and checks that every variant
produces bit-identical output (on-device checksum of the program outputs against the first variant).

    python scripts/sweep_variants.py --config 1 d4r4w8 d4r4w8p5 d4r4w8p5s ...

A variant is written d<depth>[r<rows>][v<cells>][w<warps>][k<threads per row>][p<prefetch>][s|f][x]
where a trailing ``s`` selects neighbour-only ("pair") synchronisation, ``f`` per-field mbarriers
("flags"), ``h`` two half-CTAs sharing the tile ("halves") and ``x`` direct reads of the input's neighbour rows from the TMA ring.  Generator switches
read from the environment go in front: ``SFB200_ST64=1,SFB200_SPLITBAR=1:d4r3w12p5``.
``--repeat N`` times the list N times round-robin (clock drift hits every variant alike) and reports the
median; ``--warm`` only compiles (no GPU needed: fills the program cache that travels to the GPU box).
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


ENV_KEYS = set()


def parse(text):
    """Applies the variant's environment switches (clearing those of earlier variants) and returns its options."""
    env, _, text = text.rpartition(":")
    for k in ENV_KEYS:
        os.environ.pop(k, None)
    for kv in filter(None, env.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
        ENV_KEYS.add(k)
    m = re.fullmatch(r"d(\d+)(?:r(\d+))?(?:v(\d+))?(?:w(\d+))?(?:k(\d+))?(?:p(\d+))?([sfh]?)(x?)", text)
    if not m:
        raise SystemExit("bad variant " + text)
    d, r, v, w, k, p, s, x = m.groups()
    return planner.PlanOptions(max_depth=int(d), rows_per_thread=int(r or 0), vector=int(v or 0),
                               warps=int(w or 0), threads_per_row=int(k or 0), prefetch=int(p or 0),
                               sync={"s": "pair", "f": "flags", "h": "halves"}.get(s, "cta"), direct=1 if x else 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--warm", action="store_true")
    ap.add_argument("variants", nargs="+")
    args = ap.parse_args()
    build.build_native()
    name, prog, _ = programs.baseline_config(args.config)
    path = programs.write_program(prog, name)
    nops = len(prog["program"])
    cells = float(np.prod(prog["dimensions"]))
    first = None
    if args.warm:
        for text in args.variants:
            try:
                p = CudaProgram(path, plan_options=parse(text), allocate=False)
                print("{:<40} {}  {}".format(text, "cached" if p.was_cached else "compiled",
                                             [(l.kernel, l.info.get("sync")) for l in p.lowered.launches][:2]), flush=True)
            except Exception as exc:
                print("{:<40} does not lower: {}".format(text, str(exc)[:100]), flush=True)
        return
    times = {}
    for text in args.variants * args.repeat:
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
            times.setdefault(text, []).append(ms)
            sums = []
            for oname, f in p.program.fields.items():
                if f.kind == "output":
                    n = int(np.prod(p.local_shape(oname)))
                    sums.append(p.rt.checksum(p.buffers[oname].dptr, n, f.data_type.type)[1])
            if first is None:
                first = sums
            info = p.lowered.launches[0].info
            regs = ""
            print("{:<40} {:8.3f} ms  {:.3e} upd/s  launches {}  sync {} P {} tile {} smem {}  {}".format(
                text, ms, nops * cells / (ms * 1e-3), [len(l.ops) for l in p.lowered.launches][:6],
                info.get("sync"), info.get("prefetch"), info.get("tile"), p.lowered.launches[0].smem,
                "bits==first" if sums == first else "BITS DIFFER " + str(sums) + " vs " + str(first)), flush=True)
        except Exception as exc:
            print("{:<16} failed: {}".format(text, str(exc)[:200]), flush=True)
        finally:
            p.close()
    if args.repeat > 1:
        print("---- medians")
        for text, ms in times.items():
            med = float(np.median(ms))
            print("{:<40} {:8.3f} ms  {:.3e} upd/s   {}".format(text, med, nops * cells / (med * 1e-3),
                                                              " ".join("%.3f" % m for m in ms)), flush=True)


if __name__ == "__main__":
    main()
