#!/usr/bin/env python3
"""Command line of the single-device driver.  Same positional arguments and flags as the reference's
``bin/run_program.py:12-41``; ``mode`` gains ``cuda``."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from stencilflow_b200.log_level import LogLevel  # noqa: E402
from stencilflow_b200.run_program import run_program  # noqa: E402

if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("stencil_file")
    parser.add_argument("mode", choices=["cuda", "emulation", "hardware"])
    parser.add_argument("-run-simulation", action="store_true")
    parser.add_argument("-compare-to-reference", action="store_true")
    parser.add_argument("-input-directory")
    parser.add_argument("-use-cached-sdfg", dest="use_cached_sdfg", action="store_true")
    parser.add_argument("-skip-execution", dest="skip_execution", action="store_true")
    parser.add_argument("-generate-input", action="store_true")
    parser.add_argument("-halo", type=int, default=0)
    parser.add_argument("-repetitions", type=int, default=1)
    parser.add_argument("-synthetic-reads", type=float, default=None)
    parser.add_argument("-specialize-scalars", dest="specialize_scalars", action="store_true")
    parser.add_argument("-plot", action="store_true")
    parser.add_argument("-log-level", type=int, choices=[0, 1, 2, 3], default=1)
    parser.add_argument("-print-result", dest="print_result", action="store_true")
    parser.add_argument("-xilinx", dest="xilinx", action="store_true")
    args = parser.parse_args()
    args.log_level = LogLevel(args.log_level)
    sys.exit(run_program(**vars(args)))
