#!/usr/bin/env python3
"""Ceiling of the end-to-end (host-array) call: every rank copies pinned host memory to its GPU and
back concurrently (two streams, full duplex) while all other ranks do the same; prints one JSON line
with the per-rank and aggregate GB/s per direction.

    python scripts/host_dma_ceiling.py                       # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/host_dma_ceiling.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from stencilflow_b200 import build, distributed, runtime
    build.build_native()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    comm = distributed.make_comm() if world > 1 else distributed.Comm()
    rtm = runtime.Runtime.get(int(os.environ.get("LOCAL_RANK", "0")))
    nbytes = 1 << 30
    a, pa = rtm.host_alloc((nbytes,), np.uint8)
    b, pb = rtm.host_alloc((nbytes,), np.uint8)
    a[:] = 1
    d0, d1 = rtm.malloc(nbytes), rtm.malloc(nbytes)
    s0, s1 = rtm.stream_create(), rtm.stream_create()
    results = {}
    for mode in ("h2d", "d2h", "duplex"):
        for it in range(2):
            rtm.stream_synchronize(s0)
            rtm.stream_synchronize(s1)
            comm.barrier()
            t0 = time.perf_counter()
            for _ in range(4):
                if mode in ("h2d", "duplex"):
                    rtm.h2d(d0, a, stream=s0)
                if mode in ("d2h", "duplex"):
                    rtm.d2h(b, d1, stream=s1)
            rtm.stream_synchronize(s0)
            rtm.stream_synchronize(s1)
            dt = time.perf_counter() - t0
        rate = 4 * nbytes / dt / 1e9
        rates = comm.allgather(rate)
        results[mode] = {"per_rank_gbs_per_direction": [round(r, 2) for r in rates],
                         "aggregate_gbs_per_direction": round(sum(rates), 2)}
    # the same with caller-owned numpy memory page-locked by sfb_host_register (what the reference-facing call does)
    ra, rb = np.empty(nbytes, np.uint8), np.empty(nbytes, np.uint8)
    ra[:] = 1
    rtm.host_register(ra)
    rtm.host_register(rb)
    for mode in ("h2d", "d2h", "duplex"):
        for it in range(2):
            rtm.stream_synchronize(s0)
            rtm.stream_synchronize(s1)
            comm.barrier()
            t0 = time.perf_counter()
            for _ in range(4):
                if mode in ("h2d", "duplex"):
                    rtm.h2d(d0, ra, stream=s0)
                if mode in ("d2h", "duplex"):
                    rtm.d2h(rb, d1, stream=s1)
            rtm.stream_synchronize(s0)
            rtm.stream_synchronize(s1)
            dt = time.perf_counter() - t0
        rates = comm.allgather(4 * nbytes / dt / 1e9)
        results["registered_" + mode] = {"per_rank_gbs_per_direction": [round(r, 2) for r in rates],
                                         "aggregate_gbs_per_direction": round(sum(rates), 2)}
    rtm.host_unregister(ra)
    rtm.host_unregister(rb)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "bytes_per_copy": nbytes, "host_cores": os.cpu_count(), **results}), flush=True)
    rtm.free(d0)
    rtm.free(d1)
    rtm.host_free(pa)
    rtm.host_free(pb)
    comm.close()


if __name__ == "__main__":
    main()
