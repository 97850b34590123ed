#!/usr/bin/env python3
"""Headline benchmark: stencil cell-updates/s and fraction of the HBM roofline (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C] [--impl ours|reference]

A *step* is one execution of the whole stencil program (all chained operators) over one synthetic
field.  Default workload = BASELINE.json configs[1]: Jacobi-3D, 8 chained operators, 1024^3 float32,
constant boundary.  For N > 1 (launched by torchrun, one rank per GPU) every rank owns a 1024^3 slab
of a (N*1024) x 1024 x 1024 domain and exchanges halos with its neighbours over NVLink once per
fused pass (weak scaling).

One JSON line is printed by rank 0; see DESIGN.md section "Measurement" for the definition of every key.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "stencil_cell_updates_per_s"
UNIT = "cell-updates/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        with open(self.path) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    smax.append(float(parts[2]))
                    power.append(float(parts[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "samples": len(sm),
                "power_w_max": float(max(power)), "reasons": sorted(reasons)}


VARIANT = None          # --variant: secondary form of a config (programs.VARIANTS)


def build_config(index, scale_i=1):
    from stencilflow_b200 import programs
    name, prog, halo = programs.baseline_config(index, VARIANT)
    if scale_i > 1:
        prog["dimensions"][0] *= scale_i
        name += "_x{}".format(scale_i)
    return name, prog, halo


INPUT_RANGES = {"inp": (1.0, 2.0), "coeff": (0.0, 0.05), "w": (0.9, 1.1)}   # hdiff; everything else U[0,1)


def fill_inputs(program, seed=1234, index_offsets=None):
    rt = program.rt
    for k, (name, f) in enumerate(program.program.fields.items()):
        if f.kind != "input":
            continue
        if f.is_scalar:
            program.set_scalars({name: 0.5})
            continue
        lo, hi = INPUT_RANGES.get(name, (0.0, 1.0))
        n = int(np.prod(program.local_shape(name)))
        off = (index_offsets or {}).get(name, 0)
        rt.fill_hash(program.buffers[name].dptr, n, f.data_type.type, seed + k, lo, hi, off)
    rt.stream_synchronize()


def cpu_reference_rate(index, target_seconds=12.0, threads=None):
    """Times the CPU restatement of the reference program (oracle/reference_cpp.py: OpenMP over the
    outermost loop, -O3 -march=native -ffast-math, all transients live) on a bounded sample of the
    workload.  Returns (cell-updates/s, cores, sample description)."""
    from oracle import reference_cpp
    from stencilflow_b200 import programs, synthetic
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank);
    # SFB200_REF_THREADS overrides.  The count reported is the one the OpenMP runtime really uses.
    cores = threads or int(os.environ.get("SFB200_REF_THREADS", "0")) or os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    name, prog, _ = programs.baseline_config(index, VARIANT)
    full = list(prog["dimensions"])
    nops = len(prog["program"])

    short = [d < 256 for d in full]                    # hdiff's vertical axis is kept as it is

    def sized(frac):
        return [d if sh else max(16, int(d * frac) // 8 * 8) for d, sh in zip(full, short)]

    iters = ["i", "j", "k"][3 - len(full):]

    used = [cores, 0.0]               # threads the OpenMP runtime uses, seconds of the last timed sample

    def run(dims):
        p = json.loads(json.dumps(prog))
        p["dimensions"] = dims
        ref = reference_cpp.CompiledReference(p)
        inputs = {}
        for k, (iname, cfg) in enumerate(p["inputs"].items()):
            lo, hi = INPUT_RANGES.get(iname, (0.0, 1.0))
            shape = tuple(n for it, n in zip(iters, dims) if it in cfg.get("input_dims", iters))
            inputs[iname] = synthetic.fill_hash(shape, np.dtype(cfg["data_type"]), 1234 + k, lo, hi)
        ref.allocate_transients()
        used[0] = ref.threads
        ref(**inputs)                                   # warm (page faults, OpenMP team)
        t0 = time.perf_counter()
        ref(**inputs)
        dt = time.perf_counter() - t0
        used[1] = dt
        return nops * float(np.prod(dims)) / dt, dt

    probe_dims = sized(0.125 if full[0] >= 1024 else 1.0)
    rate, _ = run(probe_dims)
    cells_full = float(np.prod(full))
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 16 << 30
    itemsize = np.dtype(next(iter(prog["inputs"].values()))["data_type"]).itemsize
    mem_cells = 0.5 * avail / ((nops + 2) * itemsize)
    want_cells = min(cells_full, rate * target_seconds / nops, mem_cells)
    frac = (want_cells / cells_full) ** (1.0 / max(1, short.count(False)))
    dims = sized(min(1.0, frac))
    if np.prod(dims) > np.prod(probe_dims):
        rate, dt = run(dims)
    else:
        dims = probe_dims
    sample = "{} at {} ({} operators, all transients live), 1 warm + 1 timed execution".format(
        name, "x".join(map(str, dims)), nops)
    cpu_reference_rate.last_sample_seconds = used[1]
    return rate, used[0], sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rates = []
    t_all0 = time.perf_counter()
    cores = os.cpu_count() or 1
    sample = ""
    times = []
    for step in range(args.warmup + args.steps):
        rate, cores, sample = cpu_reference_rate(args.config, target_seconds=6.0)
        if step >= args.warmup:
            rates.append(rate)
            times.append(cpu_reference_rate.last_sample_seconds)
        if time.perf_counter() - t_all0 > 150:
            break
    value = float(np.mean(rates)) if rates else rate
    name, prog, _ = build_config(args.config)
    dims = prog["dimensions"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(rates), "warmup": args.warmup,
        "ms_per_step": (1e3 * float(np.mean(times)) if times else None),     # one bounded sample (see "sample")
        "higher_is_better": True,
        "scaling": args.scaling or "weak", "vs_baseline": None,
        "dtype": "f64" if "f64" in name else "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, prog)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_name(index, prog):
    dims = "x".join(map(str, prog["dimensions"]))
    nops = len(prog["program"])
    dt = next(iter(prog["program"].values()))["data_type"]
    kinds = {0: "Jacobi-3D chain", 1: "Jacobi-3D chain", 2: "COSMO hdiff", 3: "Jacobi-2D chain", 4: "Jacobi-3D chain"}
    tag = {"jki": ", layout J,K,I", "w1d": ", x 1-D weight w[k] per stage"}.get(VARIANT, "")
    return "{} {} {}, {} chained operators (BASELINE.json configs[{}]{})".format(kinds[index], dims, dt, nops, index, tag)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100,
                    help="timed executions of the program (default 100: long enough for the clock "
                         "sampler to see the load and for the power cap to show)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=1, help="index into BASELINE.json configs")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="N > 1: weak = every rank owns one copy of the config's domain (default for "
                         "configs 1-3); strong = the config's domain is split (default for config 4, "
                         "the 2048^3 x 64 slab-split case of BASELINE.json)")
    ap.add_argument("--variant", default=None, choices=["jki", "w1d"],
                    help="secondary form of a config (SURVEY 8d): jki = config 2 in the reference's COSMO "
                         "layout J,K,I (1024x80x1024); w1d = config 3 with a 1-D weight w[k] in every stage")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    global VARIANT
    VARIANT = args.variant
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.scaling is None:
        args.scaling = "strong" if args.config == 4 else "weak"

    if args.impl == "reference":
        run_reference_arm(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # convenience: re-launch ourselves under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                   "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
                   "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        raise SystemExit("WORLD_SIZE={} does not match --gpus {}".format(world, args.gpus))

    from stencilflow_b200 import build
    build.build_native()
    from stencilflow_b200 import programs
    from stencilflow_b200.cuda_program import CudaProgram

    comm = None
    if world > 1:
        from stencilflow_b200 import distributed
        comm = distributed.TorchComm()

    name, prog, halo = build_config(args.config, scale_i=world if args.scaling == "weak" else 1)
    path = programs.write_program(prog, name)
    if world > 1:
        from stencilflow_b200 import distributed
        program = distributed.SlabProgram(path, comm, device=local_rank)
    else:
        program = CudaProgram(path, device=local_rank)
    peak_gbs, peak_src = load_peaks()
    fill_inputs(program, index_offsets=getattr(program, "input_index_offsets", lambda: None)())
    rtm = program.rt

    cells_total = int(np.prod(prog["dimensions"]))
    nops = len(prog["program"])
    updates_per_step = nops * cells_total                      # whole job, all ranks
    local_fraction = 1.0 / world

    def barrier():
        rtm.stream_synchronize()
        if comm is not None:
            comm.barrier()

    for _ in range(args.warmup):
        program.execute()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = program.launch_count
    e0, e1 = rtm.event_create(), rtm.event_create()
    barrier()
    rtm.event_record(e0)
    for _ in range(args.steps):
        program.execute()
    rtm.event_record(e1)
    rtm.event_synchronize(e1)
    barrier()
    ms_local = rtm.elapsed_ms(e0, e1)
    clocks = sampler.stop() if rank == 0 else None
    ms = comm.max_float(ms_local) if comm is not None else ms_local
    launches = program.launch_count - launches0
    ms_per_step = ms / args.steps
    value = updates_per_step / (ms_per_step * 1e-3)

    # roofline of the dominant kernel: algorithmic bytes of one launch / its mean duration.
    # All launches of a step are stencil passes on one stream, so the mean launch duration is
    # (event-timed step) / (launches per step).
    alg_bytes_step = program.plan.algorithmic_bytes() * (1.0 if world == 1 else local_fraction)
    lps = program.launches_per_execution
    achieved = alg_bytes_step / (ms_local / args.steps * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as f:
            rec = json.load(f).get(str(args.config), {})
            # the ncu capture only describes the kernel it was taken from
            if rec.get("kernel") == program.lowered.launches[0].kernel and world == 1:
                traffic = rec.get("dram_bytes_per_launch")
    # fusion-independent companion figure: the program's minimum off-chip volume
    # (every input read once, every output written once) over the same time
    fields = program.program.fields
    min_volume = sum(f.nbytes for f in fields.values() if f.kind in ("input", "output") and not f.is_scalar)
    min_volume *= (1.0 if world == 1 else local_fraction)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_step / lps, "launches_per_step": lps,
                "operators_per_launch": [len(l.ops) for l in program.lowered.launches],
                "min_volume_frac": min_volume / (ms_local / args.steps * 1e-3) / 1e9 / peak_gbs,
                "kernel": program.lowered.launches[0].kernel, "family": program.lowered.launches[0].family}

    # end to end through the plugin call with host buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(program, prog, args, updates_per_step, comm)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, cores, sample = cpu_reference_rate(args.config)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64" if "float64" in json.dumps(prog["program"]) else "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args.config, prog),
                       "per_gpu": "x".join(map(str, [prog["dimensions"][0] // world] + prog["dimensions"][1:])),
                       "l2": "no flush needed: every pass streams fields of {:.1f} GiB, far above the 126 MB L2".format(
                           cells_total / world * (8 if "float64" in json.dumps(prog["program"]) else 4) / 2 ** 30),
                       "plan": [{"family": l.family, "ops": len(l.ops)} for l in program.lowered.launches],
                       "input": "U[0,1) counter hash generated in HBM (seed 1234)"},
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches,
        }
        print(json.dumps(line), flush=True)
    program.close()
    if comm is not None:
        comm.close()


def measure_e2e(program, prog, args, updates_per_step, comm=None):
    """Same metric through the reference-facing call with HOST buffers: every step copies this rank's
    inputs host->device from pinned memory, runs the program and copies its outputs back
    (``CudaProgram.__call__`` on one GPU; the slab-wise equivalent on several).  Wall-clock, max over ranks."""
    rtm = program.rt
    fields = program.program.fields
    host_in, host_out, free_host = {}, {}, []
    h2d = d2h = 0
    for name, f in fields.items():
        if f.is_scalar or f.kind == "intermediate":
            continue
        shape = program.local_shape(name)
        if f.kind == "input":
            arr, hptr = rtm.host_alloc(shape, f.data_type.type)
            rtm.d2h(arr, program.buffers[name].dptr)         # reuse the synthetic field as host data
            rtm.stream_synchronize()
            host_in[name] = arr
            h2d += arr.nbytes
        else:
            arr, hptr = rtm.host_alloc(shape, f.data_type.type)
            host_out[name] = arr
            d2h += arr.nbytes
        free_host.append(hptr)
    steps = max(2, min(args.steps, 5))

    def one_step():
        # the reference-facing call: CudaProgram.__call__ / SlabProgram.__call__ with host arrays
        kw = {k + "_host": v for k, v in host_in.items()}
        kw.update({k + "_host": v for k, v in host_out.items()})
        program(**kw)

    one_step()                                               # warm
    if comm is not None:
        comm.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    rtm.stream_synchronize()
    dt = (time.perf_counter() - t0) / steps
    copied = getattr(program, "last_call_bytes", None)
    if copied:                                               # bytes the call actually moved per step
        h2d, d2h = int(copied[0]), int(copied[1])
    if comm is not None:
        dt = comm.max_float(dt)
        h2d = int(sum(comm.allgather(h2d)))
        d2h = int(sum(comm.allgather(d2h)))
    for hptr in free_host:
        rtm.host_free(hptr)
    return {"value": updates_per_step / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3, "steps": steps,
            "host_memory": "pinned (cudaHostAlloc)"}


if __name__ == "__main__":
    main()
