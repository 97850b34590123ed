"""Checks the pipelined host call at full size: PCIe copy rates, time per call, bit-identity with the plain path."""
import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import bench
from stencilflow_b200 import build, programs
from stencilflow_b200.cuda_program import CudaProgram
build.build_native()
name, prog, halo = programs.baseline_config(2)
path = programs.write_program(prog, name)
p = CudaProgram(path, device=0)
rt = p.rt
# raw copy bandwidths
n = 1 << 30
h, hp = rt.host_alloc((n,), np.uint8)
d = rt.malloc(n)
for label, fn in (("h2d", lambda: rt.h2d(d, h)), ("d2h", lambda: rt.d2h(h, d))):
    fn(); rt.stream_synchronize()
    t0 = time.perf_counter(); fn(); rt.stream_synchronize(); dt = time.perf_counter() - t0
    print(label, "%.1f GB/s" % (n / dt / 1e9))
s2 = rt.stream_create()
t0 = time.perf_counter(); rt.h2d(d, h[: n // 2]); rt.d2h(h[n // 2:], d + n // 2, stream=s2); rt.stream_synchronize(); rt.stream_synchronize(s2); dt = time.perf_counter() - t0
print("duplex 0.5+0.5 GiB: %.1f GB/s each way" % (n / 2 / dt / 1e9))
shape = tuple(prog["dimensions"])
from stencilflow_b200 import synthetic
inp, _ = rt.host_alloc(shape, np.float32); coeff, _ = rt.host_alloc(shape, np.float32)
inp[...] = synthetic.fill_hash(shape, np.float32, 1, 1.0, 2.0); coeff[...] = synthetic.fill_hash(shape, np.float32, 2, 0.0, 0.05)
outs = []
for pieces in ("0", "16"):
    os.environ["SFB200_PIPELINE_PIECES"] = pieces
    out, _ = rt.host_alloc(shape, np.float32); out[...] = 0
    rt.memset(p.buffers["inp"].dptr, 0, inp.nbytes); rt.memset(p.buffers["coeff"].dptr, 0, inp.nbytes); rt.stream_synchronize()
    p(inp_host=inp, coeff_host=coeff, out_host=out)
    print("  first call checksum", float(out.sum(dtype=np.float64)))
    t0 = time.perf_counter()
    for _ in range(5):
        p(inp_host=inp, coeff_host=coeff, out_host=out)
    print("pieces", pieces, "%.2f ms per call" % ((time.perf_counter() - t0) / 5 * 1e3), "checksum", float(out.sum(dtype=np.float64)))
    outs.append(out)
print("bit-identical:", np.array_equal(outs[0], outs[1]))
