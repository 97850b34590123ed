#!/usr/bin/env python3
"""Multi-GPU command line (counterpart of the reference's ``bin/run_distributed_program.py:56-100``).

    bin/run_distributed_program.py prog.json cuda -gpus 4 -compare-to-reference [-halo H]

Started once, it starts one process per GPU itself (plain subprocesses that find each other over TCP on
127.0.0.1 -- ``distributed.SocketComm``; no MPI, no torch); started by a launcher that sets RANK / WORLD_SIZE
(torchrun, mpirun-style wrappers) it simply runs its rank."""
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
        import socket
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        procs = []
        for rank in range(n):
            env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(n), MASTER_ADDR="127.0.0.1",
                       MASTER_PORT=str(port), SFB200_COMM_PORT=str(port), SFB200_COMM="socket")
            procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env))
        codes = [p.wait() for p in procs]
        sys.exit(max(codes, key=abs))
    from stencilflow_b200.run_distributed import run_distributed_program
    ret = run_distributed_program(args.stencil_file, args.mode, compare_to_reference=args.compare_to_reference,
                                  input_directory=args.input_directory, halo=args.halo,
                                  repetitions=args.repetitions, log_level=args.log_level)
    sys.exit(ret or 0)
