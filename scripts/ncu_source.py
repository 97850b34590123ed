#!/usr/bin/env python3
"""Per-opcode aggregation of an ncu source page: executed instructions, stall samples, shared-memory
wavefronts (actual / ideal).   usage: scripts/ncu_source.py report.ncu-rep [--top N]"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
    total_samples = 0
    stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
    stall_tot = collections.Counter()
    per_line = []
    for r in rows:
        src = re.sub(r"^@!?U?P\d+\s+", "", r["Source"].strip())
        opc = src.split()[0].split(".")[0] if src else "?"
        if opc in ("LDS", "STS", "LDG", "STG", "ST", "LD"):
            opc = src.split()[0]
        a = agg[opc]
        ex = int(r["Instructions Executed"] or 0)
        smp = int(r["# Samples"] or 0)
        a[0] += ex
        a[1] += smp
        a[2] += int(r["L1 Wavefronts Shared"] or 0)
        a[3] += int(r["L1 Wavefronts Shared Ideal"] or 0)
        a[4] += 1
        total_samples += smp
        for c in stall_cols:
            stall_tot[c] += int(r[c] or 0)
        per_line.append((smp, r["Source"].strip(), {c: int(r[c] or 0) for c in stall_cols if int(r[c] or 0)}))
    tot_ex = sum(a[0] for a in agg.values())
    print("total warp instructions {:,}  samples {:,}".format(tot_ex, total_samples))
    print("{:<14}{:>14}{:>8}{:>10}{:>8}{:>14}{:>14}".format("opcode", "executed", "%", "samples", "%", "smem wavefr", "ideal"))
    for opc, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("{:<14}{:>14,}{:>8.1f}{:>10,}{:>8.1f}{:>14,}{:>14,}".format(
            opc, a[0], 100.0 * a[0] / max(1, tot_ex), a[1], 100.0 * a[1] / max(1, total_samples), a[2], a[3]))
    print("stall totals:", ", ".join("{} {:.1f}%".format(k[6:], 100.0 * v / max(1, total_samples))
                                     for k, v in stall_tot.most_common(10)))
    if "--lines" in sys.argv:
        for smp, src, st in sorted(per_line, key=lambda x: -x[0])[:40]:
            print("{:>7} {:<70} {}".format(smp, src[:70], st))


if __name__ == "__main__":
    main()
