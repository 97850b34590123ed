#!/usr/bin/env python3
"""Condenses an .ncu-rep (ncu --set full) into the few numbers DESIGN.md / bench.py cite.
usage: scripts/ncu_summary.py report.ncu-rep [algorithmic_bytes_per_launch]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    rep = sys.argv[1]
    alg = float(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        rec = dict(zip(hdr, row))
        print("kernel:", rec.get("Kernel Name"), " grid", rec.get("Grid Size"), " block", rec.get("Block Size"))
        for k in KEYS:
            if k in rec:
                print("  {:<78} {:>18} {}".format(k, rec[k], units[hdr.index(k)]))
        stalls = []
        prefix, suffix = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"
        for h in hdr:
            if h.startswith(prefix) and h.endswith(suffix):
                try:
                    stalls.append((float(rec[h]), h[len(prefix):-len(suffix)]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  warp stall reasons (warps per issued instruction):",
              ", ".join("{} {:.2f}".format(n, v) for v, n in stalls[:8]))
        try:
            rd = float(rec["dram__bytes_read.sum"]) * SCALE[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(rec["dram__bytes_write.sum"]) * SCALE[units[hdr.index("dram__bytes_write.sum")]]
            print("  DRAM traffic per launch: {:.4e} B (read {:.4e} + write {:.4e})".format(rd + wr, rd, wr))
            if alg:
                print("  algorithmic bytes per launch: {:.4e} B  -> traffic / algorithmic = {:.3f}".format(
                    alg, (rd + wr) / alg))
        except (KeyError, ValueError):
            pass


if __name__ == "__main__":
    main()
