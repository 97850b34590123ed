#!/usr/bin/env python3
"""Regenerates the JSON programs under tests/programs/.

* ``ref_*``: the programs of the reference's own test-suite (``/root/reference/test/stencils/*.json``,
  listed by ``test/test_stencilflow.py:191-216`` and ``test/test_distributed_program.sh``), re-serialised.
  The reference tree only exists in the build container, so the files are committed.
* everything else: programs written for this repository to cover what the reference programs do not
  (copy/shrink boundaries, hdiff, ternaries and math calls, 1-D programs, box taps, mixed types).
Run from the repository root:  python tests/programs/make_programs.py
"""
import glob
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/test/stencils"


def dump(name, prog):
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(prog, f, indent=1, sort_keys=True)
        f.write("\n")


def jacobi3d(n, steps, dtype="float32", bc=None, coeff="0.16666666", data="constant:1.0", shape=None):
    shape = shape or [n, n, n]
    bc = bc or {"type": "constant", "value": 0.0}
    prog = {"inputs": {"a": {"data": data, "data_type": dtype}}, "outputs": ["b%d" % (steps - 1)],
            "dimensions": shape, "program": {}}
    prev = "a"
    for s in range(steps):
        name = "b%d" % s
        prog["program"][name] = {
            "computation_string": "{n} = {c} * ({p}[i-1,j,k] + {p}[i+1,j,k] + {p}[i,j-1,k] + {p}[i,j+1,k] + "
                                  "{p}[i,j,k-1] + {p}[i,j,k+1])".format(n=name, c=coeff, p=prev),
            "boundary_conditions": {prev: dict(bc)}, "data_type": dtype}
        prev = name
    return prog


def jacobi2d(nj, nk, steps, dtype="float64", bc=None, data="constant:1.0"):
    bc = bc or {"type": "shrink"}
    prog = {"inputs": {"a": {"data": data, "data_type": dtype}}, "outputs": ["b%d" % (steps - 1)],
            "dimensions": [nj, nk], "program": {}}
    prev = "a"
    for s in range(steps):
        name = "b%d" % s
        prog["program"][name] = {
            "computation_string": "{n} = 0.25 * ({p}[j-1,k] + {p}[j+1,k] + {p}[j,k-1] + {p}[j,k+1])".format(
                n=name, p=prev),
            "boundary_conditions": {prev: dict(bc)}, "data_type": dtype}
        prev = name
    return prog


def hdiff(ni, nj, nk, dtype="float32"):
    """COSMO horizontal diffusion (laplacian -> flux limiter in i and j -> output), SURVEY section 8d config 3."""
    sh = {"type": "shrink"}
    return {
        "inputs": {"inp": {"data": "constant:1.0", "data_type": dtype},
                   "coeff": {"data": "constant:0.025", "data_type": dtype}},
        "outputs": ["out"], "dimensions": [ni, nj, nk],
        "program": {
            "lap": {"computation_string": "lap = 4.0*inp[i,j,k] - (inp[i+1,j,k] + inp[i-1,j,k] + inp[i,j+1,k] + inp[i,j-1,k])",
                    "boundary_conditions": {"inp": sh}, "data_type": dtype},
            "flx": {"computation_string": "d = lap[i+1,j,k] - lap[i,j,k]; flx = 0.0 if d*(inp[i+1,j,k] - inp[i,j,k]) > 0.0 else d",
                    "boundary_conditions": {"lap": sh, "inp": sh}, "data_type": dtype},
            "fly": {"computation_string": "d = lap[i,j+1,k] - lap[i,j,k]; fly = 0.0 if d*(inp[i,j+1,k] - inp[i,j,k]) > 0.0 else d",
                    "boundary_conditions": {"lap": sh, "inp": sh}, "data_type": dtype},
            "out": {"computation_string": "out = inp[i,j,k] - coeff[i,j,k]*(flx[i,j,k] - flx[i-1,j,k] + fly[i,j,k] - fly[i,j-1,k])",
                    "boundary_conditions": {"inp": sh, "coeff": sh, "flx": sh, "fly": sh}, "data_type": dtype},
        }}


def main():
    # all-zero raw float32 inputs used by ref_jacobi2d_128x128 / ref_jacobi3d_32x32x32
    # (reference test/stencils/data/zeros_*.dat)
    import numpy as np
    os.makedirs(os.path.join(HERE, "data"), exist_ok=True)
    np.zeros(128 * 128, dtype=np.float32).tofile(os.path.join(HERE, "data", "zeros_128x128_fp32.dat"))
    np.zeros(32 ** 3, dtype=np.float32).tofile(os.path.join(HERE, "data", "zeros_32x32x32_fp32.dat"))
    if os.path.isdir(REF):
        for path in sorted(glob.glob(os.path.join(REF, "*.json"))):
            with open(path) as f:
                dump("ref_" + os.path.splitext(os.path.basename(path))[0], json.load(f))
    dump("jacobi3d_16x24x32_5itr_const1", jacobi3d(0, 5, bc={"type": "constant", "value": 1.0}, shape=[16, 24, 32]))
    dump("jacobi3d_24x20x40_4itr_shrink_f64", jacobi3d(0, 4, dtype="float64", bc={"type": "shrink"},
                                                     coeff="0.16666666666666666", shape=[24, 20, 40]))
    dump("jacobi3d_12x12x16_3itr_copy", jacobi3d(0, 3, bc={"type": "copy"}, shape=[12, 12, 16]))
    p = jacobi3d(0, 3, bc={"type": "copy"}, shape=[12, 12, 16])
    for s, (name, entry) in enumerate(p["program"].items()):
        prev = "a" if s == 0 else "b%d" % (s - 1)
        entry["computation_string"] = entry["computation_string"].replace(
            "({}[i-1,j,k]".format(prev), "({p}[i,j,k] + {p}[i-1,j,k]".format(p=prev))
    dump("jacobi3d_12x12x16_3itr_copy", p)
    dump("jacobi2d_96x128_6itr_shrink_f64", jacobi2d(96, 128, 6))
    dump("jacobi2d_64x64_4itr_const_f32", jacobi2d(64, 64, 4, dtype="float32", bc={"type": "constant", "value": 0.5}))
    dump("hdiff_24x28x16", hdiff(24, 28, 16))
    dump("hdiff_16x20x8_f64", hdiff(16, 20, 8, "float64"))
    dump("box3d_10x12x16", {
        "inputs": {"a": {"data": "constant:1.0", "data_type": "float32"}},
        "outputs": ["c"], "dimensions": [10, 12, 16],
        "program": {
            "b": {"computation_string": "b = 0.125 * (a[i-1,j-1,k-1] + a[i-1,j+1,k+1] + a[i+1,j-1,k+1] + a[i+1,j+1,k-1] + "
                                        "a[i,j,k-2] + a[i,j,k+2] + a[i,j-2,k] + a[i+2,j,k])",
                  "boundary_conditions": {"a": {"type": "constant", "value": 2.0}}, "data_type": "float32"},
            "c": {"computation_string": "c = b[i,j,k] + 0.5 * b[i-1,j,k+3]",
                  "boundary_conditions": {"b": {"type": "constant", "value": 1.0}}, "data_type": "float32"}}})
    # within the envelope of the reference's dataflow simulator (3-D, full-dimensional list inputs,
    # constant boundaries, no and/or; calculator.py knows sin/cos/tan/sinh/cosh): these two run through
    # the reference itself, tests/golden/make_reference_sim_golden.py
    dump("trig3d_8x10x12_f64", {
        "inputs": {"a": {"data": "constant:0.75", "data_type": "float64"},
                   "b": {"data": "constant:0.25", "data_type": "float64"}},
        "outputs": ["v"], "dimensions": [8, 10, 12],
        "program": {
            "u": {"computation_string": "u = sin(a[i,j,k]) * b[i-1,j,k+1] + cos(b[i,j+1,k]) / (1.5 + a[i+1,j,k-1])",
                  "boundary_conditions": {"a": {"type": "constant", "value": 0.25},
                                          "b": {"type": "constant", "value": 0.5}}, "data_type": "float64"},
            "v": {"computation_string": "v = u[i,j,k] if u[i,j-1,k] > a[i,j,k] else (u[i,j,k+2] + 0.5 * u[i-2,j,k])",
                  "boundary_conditions": {"u": {"type": "constant", "value": 1.0},
                                          "a": {"type": "constant", "value": 0.25}}, "data_type": "float64"}}})
    hc = hdiff(10, 12, 8, "float64")
    for op in hc["program"].values():
        op["boundary_conditions"] = {k: {"type": "constant", "value": 0.5} for k in op["boundary_conditions"]}
    dump("hdiff_const_10x12x8_f64", hc)
    half = {"type": "constant", "value": 0.5}
    dump("multistmt3d_6x8x10_f64", {
        "inputs": {"a": {"data": "constant:1.0", "data_type": "float64"}}, "outputs": ["q"], "dimensions": [6, 8, 10],
        "program": {
            "p": {"computation_string": "d = a[i+1,j,k] - a[i,j,k]; p = 0.0 if d*(a[i,j+1,k] - a[i,j,k]) > 0.0 else d",
                  "boundary_conditions": {"a": dict(half)}, "data_type": "float64"},
            "q": {"computation_string": "q = tan(p[i,j,k-1]) + sinh(p[i,j,k]) / cosh(p[i-1,j,k]) if p[i,j+1,k] <= 0.25 "
                                        "else (p[i,j,k] if p[i,j,k] >= 0.1 else 2.0*p[i,j,k])",
                  "boundary_conditions": {"p": dict(half)}, "data_type": "float64"}}})
    dump("diamond3d_12x10x16", {
        "inputs": {"a": {"data": "constant:1.0", "data_type": "float32"},
                   "c": {"data": "constant:0.5", "data_type": "float32"}},
        "outputs": ["e"], "dimensions": [12, 10, 16],
        "program": {
            "b": {"computation_string": "b = 0.5 * (a[i-1,j,k] + a[i+1,j,k]) + c[i,j,k]",
                  "boundary_conditions": {"a": {"type": "constant", "value": 0.0}}, "data_type": "float32"},
            "l": {"computation_string": "l = 0.5 * (b[i,j-1,k] + b[i,j+1,k])",
                  "boundary_conditions": {"b": {"type": "constant", "value": 0.0}}, "data_type": "float32"},
            "r": {"computation_string": "r = 0.25 * (b[i,j,k-1] + b[i,j,k+1]) + a[i,j,k] * c[i,j+1,k]",
                  "boundary_conditions": {"b": {"type": "constant", "value": 0.0},
                                          "c": {"type": "constant", "value": 2.0}}, "data_type": "float32"},
            "e": {"computation_string": "e = l[i,j,k] + r[i+1,j,k] + 0.5 * r[i-1,j,k]",
                  "boundary_conditions": {"r": {"type": "constant", "value": 0.0}}, "data_type": "float32"}}})
    dump("math_ops_8x8x8", {
        "inputs": {"a": {"data": "constant:0.75", "data_type": "float32"},
                   "b": {"data": "constant:0.25", "data_type": "float32"},
                   "w": {"data": 1.5, "data_type": "float32", "input_dims": []}},
        "outputs": ["r"], "dimensions": [8, 8, 8],
        "constants": {"c0": {"value": 0.5, "data_type": "float32"}},
        "program": {
            "t": {"computation_string": "s = sqrt(a[i,j,k] + b[i,j,k+1]); t = max(s, c0) + min(a[i,j-1,k], b[i,j,k]) * w",
                  "boundary_conditions": {"a": {"type": "constant", "value": 1.0}, "b": {"type": "constant", "value": 4.0}},
                  "data_type": "float32"},
            "r": {"computation_string": "r = (t[i,j,k] if (t[i,j,k] > 0.9 and a[i,j,k] < 1.0) or t[i-1,j,k] <= 0.0 else 3.0*t[i,j,k]) / 2.0 + fabs(cos(t[i,j,k-1]))",
                  "boundary_conditions": {"t": {"type": "constant", "value": 0.0}, "a": {"type": "constant", "value": 0.0}},
                  "data_type": "float32"}}})
    dump("smooth1d_256", {
        "inputs": {"x": {"data": "constant:2.0", "data_type": "float64"}},
        "outputs": ["z"], "dimensions": [256],
        "program": {
            "y": {"computation_string": "y = 0.25*x[k-1] + 0.5*x[k] + 0.25*x[k+1]",
                  "boundary_conditions": {"x": {"type": "constant", "value": 0.0}}, "data_type": "float64"},
            "z": {"computation_string": "z = y[k+2] + 0.5*y[k-2]",
                  "boundary_conditions": {"y": {"type": "copy"}}, "data_type": "float64"}}})
    p = json.loads(json.dumps(p))
    dump("lowdim2d_32x64", {
        "inputs": {"a": {"data": "constant:1.0", "data_type": "float64"},
                   "w": {"data": "constant:0.5", "data_type": "float64", "input_dims": ["k"]},
                   "h": {"data": "constant:2.0", "data_type": "float32", "input_dims": ["j"]}},
        "outputs": ["b"], "dimensions": [32, 64],
        "program": {
            "b": {"computation_string": "b = w[k] * (a[j-1,k] + a[j+1,k]) + w[k+1] * h[j] * a[j,k-1]",
                  "boundary_conditions": {"a": {"type": "constant", "value": 3.0}, "w": {"type": "constant", "value": 0.0},
                                          "h": {"type": "constant", "value": 0.0}}, "data_type": "float64"}}})
    # lower-dimensional inputs inside fusable chains (every subset of the iterators, with offsets):
    # the streamed kernels read them straight from global memory
    c1 = {"type": "constant", "value": 1.0}
    sh = {"type": "shrink"}
    low = {
        "inputs": {"a": {"data": "constant:1.0", "data_type": "float32"},
                   "wk": {"data": "constant:0.5", "data_type": "float32", "input_dims": ["k"]},
                   "wj": {"data": "constant:0.5", "data_type": "float32", "input_dims": ["j"]},
                   "wi": {"data": "constant:0.5", "data_type": "float32", "input_dims": ["i"]},
                   "pjk": {"data": "constant:0.5", "data_type": "float32", "input_dims": ["j", "k"]},
                   "pik": {"data": "constant:0.5", "data_type": "float32", "input_dims": ["i", "k"]},
                   "pij": {"data": "constant:0.5", "data_type": "float32", "input_dims": ["i", "j"]}},
        "outputs": ["b2"], "dimensions": [20, 24, 48],
        "program": {
            "b0": {"computation_string": "b0 = 0.2*wk[k]*(a[i-1,j,k] + a[i+1,j,k] + a[i,j-1,k] + a[i,j+1,k]) + "
                                         "wj[j+1]*a[i,j,k-1] + wi[i-1]*a[i,j,k+1] + pjk[j,k+1]",
                   "boundary_conditions": {"a": c1, "wk": c1, "wj": c1, "wi": c1, "pjk": c1}, "data_type": "float32"},
            "b1": {"computation_string": "b1 = 0.25*(b0[i-1,j,k] + b0[i+1,j,k] + b0[i,j-1,k] + b0[i,j,k+1]) * "
                                         "pik[i+1,k-1] + pij[i,j] - wk[k-3]",
                   "boundary_conditions": {"b0": c1, "pik": c1, "pij": c1, "wk": c1}, "data_type": "float32"},
            "b2": {"computation_string": "b2 = (b1[i,j,k] + b1[i,j+1,k] + b1[i,j,k-1]) * pij[i-1,j+1] + pjk[j-1,k] * wi[i]",
                   "boundary_conditions": {"b1": c1, "pij": c1, "pjk": c1, "wi": c1}, "data_type": "float32"}}}
    dump("lowdim3d_20x24x48_3st_f32", low)
    low64 = json.loads(json.dumps(low).replace("float32", "float64"))
    for op in low64["program"].values():
        op["boundary_conditions"] = {k: dict(sh) for k in op["boundary_conditions"]}
    dump("lowdim3d_20x24x48_3st_shrink_f64", low64)
    # BASELINE configs[3], secondary variant: the 2-D chain multiplied by 1-D weights
    for name, shape, dtype, bc in (("jacobi2d_96x128_6itr_w1d_shrink_f64", [96, 128], "float64", sh),
                                   ("jacobi2d_64x256_6itr_w1d_const_f32", [64, 256], "float32",
                                    {"type": "constant", "value": 0.5})):
        p2 = jacobi2d(shape[0], shape[1], 6, dtype=dtype, bc=bc)
        p2["inputs"]["w"] = {"data": "constant:0.9", "data_type": dtype, "input_dims": ["k"]}
        p2["inputs"]["u"] = {"data": "constant:1.1", "data_type": dtype, "input_dims": ["j"]}
        for n, op in enumerate(p2["program"].values()):
            if n % 2 == 0:
                op["computation_string"] = op["computation_string"].replace("0.25 *", "0.25 * w[k] *")
                op["boundary_conditions"]["w"] = dict(bc)
            else:
                op["computation_string"] = op["computation_string"].replace("0.25 *", "0.25 * w[k+1] * u[j-1] *")
                op["boundary_conditions"]["w"] = dict(bc)
                op["boundary_conditions"]["u"] = dict(bc)
        dump(name, p2)
    dump("fork_join_20x16x24", {
        "inputs": {"a": {"data": "constant:1.0", "data_type": "float32"}},
        "outputs": ["e", "c"], "dimensions": [20, 16, 24],
        "program": {
            "b": {"computation_string": "b = 0.5 * (a[i-1,j,k] + a[i+1,j,k])",
                  "boundary_conditions": {"a": {"type": "constant", "value": 0.0}}, "data_type": "float32"},
            "c": {"computation_string": "c = 0.5 * (b[i,j-1,k] + b[i,j+1,k])",
                  "boundary_conditions": {"b": {"type": "constant", "value": 0.0}}, "data_type": "float32"},
            "d": {"computation_string": "d = 0.5 * (b[i,j,k-1] + b[i,j,k+1]) + a[i,j,k]",
                  "boundary_conditions": {"b": {"type": "constant", "value": 0.0}}, "data_type": "float32"},
            "e": {"computation_string": "e = c[i,j,k] + d[i+1,j,k] + 0.5*d[i-1,j,k]",
                  "boundary_conditions": {"d": {"type": "constant", "value": 0.0}}, "data_type": "float32"}}})
    # programs written by the synthetic generator (the reference's bin/synthesize.py conventions):
    # box taps with extra off-chip fields, per-tap 0-D coefficients, the Rodinia hotspot formulas in
    # 3-D and 2-D, and a chain with a fork and a join
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from stencilflow_b200 import synthesize as syn
    dump("synth_box_12x10x16_3st", syn.synthesize("float32", 3, 0.5, 12, 10, 16, 1, 1, 1, stencil_shape="box"))
    dump("synth_diffusion_10x12x16_4st", syn.synthesize("float32", 4, 0, 10, 12, 16, 1, 1, 1, stencil_shape="diffusion"))
    dump("synth_hotspot3d_12x12x16_4st", syn.synthesize("float32", 4, 0.5, 12, 12, 16, 1, 1, 1, stencil_shape="hotspot"))
    dump("synth_hotspot2d_48x64_4st_f64", syn.synthesize("float64", 4, 0, 48, 64, 0, 1, 1, 0, stencil_shape="hotspot"))
    dump("synth_fork_16x12x16_5st", syn.synthesize("float32", 5, 0.3, 16, 12, 16, 1, 1, 1, fork_frequency=0.5,
                                                  fork_length_left=1, fork_length_right=2))


if __name__ == "__main__":
    main()
