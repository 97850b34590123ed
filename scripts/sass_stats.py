#!/usr/bin/env python3
"""CPU-side check of a generated kernel: compile a program (NVRTC, no GPU needed), disassemble the
cubin and print, per kernel, registers / spills and the instruction mix of its hot loop (the body of
the largest backward branch), normalised per cell update.

    python scripts/sass_stats.py programs/jacobi3d_1024_8itr_f32.json [--mix]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from stencilflow_b200.cuda_program import CudaProgram  # noqa: E402

INSTR = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;?\s*/\*")


def main():
    path = sys.argv[1]
    mix = "--mix" in sys.argv
    p = CudaProgram(path, allocate=False)
    cubin = os.path.join(p.cache_dir, "kernel.cubin")
    res = subprocess.run(["cuobjdump", "-res-usage", cubin], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    usage = {}
    cur = None
    for line in res.stdout.splitlines():
        m = re.search(r"Function (\w+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = line.strip()
    sass = subprocess.run(["cuobjdump", "-sass", cubin], stdout=subprocess.PIPE, text=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\w+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = INSTR.match(line)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), m.group(2)))
    launches = {l.kernel: l for l in p.lowered.launches}
    for name, ins in funcs.items():
        l = launches.get(name)
        print("==", name, usage.get(name, ""))
        if l is not None:
            print("   ops/pass", len(l.ops), {k: v for k, v in l.info.items() if k in ("V", "R", "warps", "tile", "block_out", "halo", "unroll", "packed", "window_registers", "lags", "windows")})
        best = None
        for addr, text in ins:
            m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", text)
            if m:
                target = int(m.group(1), 16)
                if target < addr and (best is None or addr - target > best[1] - best[0]):
                    best = (target, addr)
        if best is None:
            print("   no loop")
            continue
        body = [t for a, t in ins if best[0] <= a <= best[1]]
        counts = collections.Counter()
        for t in body:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            counts[t.split()[0].split(".")[0]] += 1
        total = len(body)
        line = "   loop {} instr".format(total)
        if l is not None and l.family == "streamed":
            cells = l.info["R"] * l.info["V"] * len(l.ops) * l.info.get("unroll", 1)
            line += " = {:.2f} per computed cell update ({} updates per thread per trip)".format(total / cells, cells)
        print(line)
        fp = sum(counts[k] for k in ("FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2", "DADD", "DMUL", "DFMA"))
        mv = sum(counts[k] for k in ("MOV", "IMAD", "PRMT", "SEL", "FSEL"))
        mem = sum(counts[k] for k in ("LDS", "STS", "STG", "LDG", "SHFL"))
        print("   fp {}  mov/imad/sel {}  shfl/lds/sts/stg {}  other {}".format(fp, mv, mem, total - fp - mv - mem))
        if mix:
            print("   " + "  ".join("{} {}".format(k, v) for k, v in counts.most_common(30)))


if __name__ == "__main__":
    main()
