#!/usr/bin/env python3
"""Tracked SASS evidence for a BASELINE config: compiles the planned program (NVRTC, no GPU needed),
disassembles the cubin (cuobjdump -sass) and prints, per kernel, the resource usage and the opcode
histogram of the whole kernel (modifiers kept for the memory / TMA / barrier opcodes).

    python scripts/sass_histogram.py 1 > profiles/r02_config1_sass.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from stencilflow_b200 import programs  # noqa: E402
from stencilflow_b200.cuda_program import CudaProgram  # noqa: E402

KEEP_MODIFIERS = ("LDS", "STS", "LDG", "STG", "UTMALDG", "UTMASTG", "SYNCS", "BAR", "SHFL", "LDL", "STL")


def main():
    index = int(sys.argv[1])
    name, prog, _ = programs.baseline_config(index)
    path = programs.write_program(prog, name)
    p = CudaProgram(path, allocate=False)
    cubin = os.path.join(p.cache_dir, "kernel.cubin")
    print("program {}  plan {}".format(name, [(l.family, len(l.ops), l.kernel) for l in p.lowered.launches][:4]))
    for l in p.lowered.launches[:1]:
        print("geometry", {k: v for k, v in l.info.items() if not callable(v)})
    res = subprocess.run(["cuobjdump", "-res-usage", cubin], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    print(res.strip())
    sass = subprocess.run(["cuobjdump", "-sass", cubin], stdout=subprocess.PIPE, text=True).stdout
    cur, hist = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\w+)", line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"^\s+/\*[0-9a-f]{4,5}\*/\s+(.*?)\s*;", line)
        if m and cur:
            text = re.sub(r"^@!?U?P\d+\s+", "", m.group(1))
            opc = text.split()[0]
            base = opc.split(".")[0]
            hist[cur][opc if base in KEEP_MODIFIERS else base] += 1
    for fn, h in hist.items():
        total = sum(h.values())
        print("\n== {}: {} SASS instructions".format(fn, total))
        for opc, n in h.most_common():
            print("  {:<34}{:>7}{:>7.1f} %".format(opc, n, 100.0 * n / total))


if __name__ == "__main__":
    main()
