#!/usr/bin/env python3
"""Static performance model of a stencil program.

Same command line and the same three sections as the reference's ``bin/report.py:11-57`` (operation
counts, runtime bounds and memory requirements of the fully pipelined dataflow design at a given
clock), followed by what replaces them on a B200: the pass plan the planner chose, the bytes every
pass moves and the resulting HBM-roofline floor.

    report.py program.json frequency_mhz [-hbm-gbs GB/s] [-no-plan]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from stencilflow_b200.kernel_chain_graph import KernelChainGraph  # noqa: E402
from stencilflow_b200.log_level import LogLevel  # noqa: E402


def dataflow_model(chain, mhz, out=print):
    """The reference's analytical model (``bin/report.py:17-57``), line for line."""
    hz = mhz * 1e6
    ops = chain.operation_count()
    width = chain.vectorization or 1
    cycles = chain.runtime_lower_bound()
    volume = chain.minimum_communication_volume()
    times = "{}*".format(width) if width != 1 else ""
    per_cycle = sum(c for c, _ in ops.values())
    total = sum(t for _, t in ops.values())
    out("======== Compute performance ============================")
    for name, (count, count_total) in ops.items():
        if width == 1:
            out("Operations per cycle multiplied by vector length {}".format(width))
        out("{}: {}{} per cycle ({} for program)".format(name, times, count, count_total))
    out("Total: {}{} per cycle ({} for program)".format(times, per_cycle, total))
    out("Upper bound on performance at {} MHz: {} GOp/s".format(mhz, 1e-9 * total / cycles * hz))
    out("Peak performance at {} MHz: {} GOp/s".format(mhz, 1e-9 * (width * per_cycle * hz)))
    out("======== Runtime ========================================")
    out("Lower bound on runtime: {} cycles ({} seconds at {} MHz)".format(cycles, cycles / hz, mhz))
    out("Peak runtime: {} cycles ({} seconds at {} MHz)".format(
        total // (width * per_cycle), total / (per_cycle * width) / hz, mhz))
    out("======== Memory performance =============================")
    ndim = len(chain.dimensions)
    operands = sum(width if len(cfg["input_dims"]) == ndim else 1 for cfg in chain.inputs.values())
    out("Number of memory accesses per cycle: {} operands".format(operands + width * len(chain.outputs)))
    out("Lower bound communication volume: {} MB".format(1e-6 * volume))
    out("Required bandwidth: {} GB/s".format(1e-9 * volume / (cycles / hz)))
    return {"operations": total, "cycles": cycles, "volume": volume}


def gpu_plan(path, hbm_gbs, out=print):
    """Pass plan of the CUDA backend for the same program: what stays on chip and what crosses HBM."""
    from stencilflow_b200.cuda_program import CudaProgram
    program = CudaProgram(path, allocate=False)
    plan = program.plan.describe()
    out("======== B200 pass plan =================================")
    updates = plan["cell_updates"]
    total_bytes = 0
    for n, p in enumerate(plan["passes"]):
        info = p["info"]
        shape = ""
        if p["family"] == "streamed":
            shape = " tile {}x{} ({} threads, {} B shared memory, TMA ring {}), tile efficiency {:.2f}".format(
                info["tile"][0], info["tile"][1], p["block"][0], p["smem"], info["prefetch"] + 1,
                info["tile_efficiency"])
        out("pass {}: {} [{} operator(s): {}] reads {} writes {}: {} bytes{}".format(
            n, p["family"], len(p["ops"]), ", ".join(p["ops"]), p["reads"], p["writes"], p["algorithmic_bytes"], shape))
        total_bytes += p["algorithmic_bytes"]
    floor = total_bytes / (hbm_gbs * 1e9)
    out("Off-chip volume of the plan: {} MB in {} pass(es) (fully fused lower bound: {} MB)".format(
        1e-6 * total_bytes, len(plan["passes"]), 1e-6 * program.chain.minimum_communication_volume()))
    out("HBM roofline floor at {} GB/s: {} seconds ({} cell updates/s)".format(
        hbm_gbs, floor, updates / floor if floor else float("inf")))
    if plan.get("tuned_from"):
        out("Plan taken from the measured-plan table: {}".format(json.dumps(plan["tuned_from"])))
    return plan


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("input_file")
    ap.add_argument("frequency", type=float)
    ap.add_argument("-hbm-gbs", dest="hbm_gbs", type=float, default=None,
                    help="HBM bandwidth for the roofline floor (default: MEASURED_PEAKS.json, else 6650)")
    ap.add_argument("-no-plan", dest="no_plan", action="store_true", help="only the reference's sections")
    args = ap.parse_args(argv)
    chain = KernelChainGraph(path=args.input_file, log_level=LogLevel.NO_LOG)
    dataflow_model(chain, args.frequency)
    if not args.no_plan:
        hbm = args.hbm_gbs
        if hbm is None:
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                    hbm = float(json.load(f)["hbm_gbs"])
            except (OSError, KeyError, ValueError):
                hbm = 6650.0
        gpu_plan(args.input_file, hbm)
    return 0


if __name__ == "__main__":
    sys.exit(main())
