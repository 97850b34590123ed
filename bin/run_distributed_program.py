#!/usr/bin/env python3
"""Multi-GPU command line (counterpart of the reference's ``bin/run_distributed_program.py:56-100``).

    bin/run_distributed_program.py prog.json cuda -gpus 4 -compare-to-reference [-halo H]

Started once, it launches one process per GPU (torch.distributed.run on 127.0.0.1); started under
torchrun/mpirun-style launchers that set RANK/WORLD_SIZE it simply runs its rank."""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("stencil_file", help="JSON description of the stencil")
    parser.add_argument("mode", choices=["cuda", "emulation", "hardware"], help="Execution mode")
    parser.add_argument("-gpus", type=int, default=0, help="number of GPUs (default: WORLD_SIZE or all)")
    parser.add_argument("-compare-to-reference", action="store_true")
    parser.add_argument("-input-directory")
    parser.add_argument("-halo", type=int, default=0)
    parser.add_argument("-repetitions", type=int, default=1)
    parser.add_argument("-log-level", type=int, choices=[0, 1, 2, 3], default=1)
    args = parser.parse_args()
    if "RANK" not in os.environ:
        n = args.gpus
        if n <= 0:
            from stencilflow_b200 import runtime
            import ctypes
            c = ctypes.c_int(0)
            runtime.load_library().sfb_device_count(ctypes.byref(c))
            n = max(1, c.value)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
               "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 500),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    from stencilflow_b200.run_distributed import run_distributed_program
    ret = run_distributed_program(args.stencil_file, args.mode, compare_to_reference=args.compare_to_reference,
                                  input_directory=args.input_directory, halo=args.halo,
                                  repetitions=args.repetitions, log_level=args.log_level)
    sys.exit(ret or 0)
