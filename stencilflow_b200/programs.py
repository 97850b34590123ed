"""Builders for the benchmark programs named in BASELINE.json (JSON dicts in StencilFlow's format).

``jacobi3d_chain`` / ``jacobi2d_chain`` emit what the reference's ``bin/synthesize.py`` emits for the
``cross`` shape (``bin/synthesize.py:95-168``: ``b<t> = <1/n>*(sum of the 2*ndim face neighbours)``,
input ``a``, constant boundary 0), with the boundary type selectable; ``hdiff`` is the COSMO horizontal
diffusion program of SURVEY section 8d (config 3), which the reference does not ship.
"""

import json
import os


def jacobi3d_chain(shape, steps, dtype="float32", boundary=None, coeff=None, data="constant:1.0",
                   vectorization=1):
    boundary = boundary or {"type": "constant", "value": 0.0}
    coeff = coeff if coeff is not None else "0.16666666"
    prog = {"inputs": {"a": {"data": data, "data_type": dtype}}, "outputs": ["b%d" % (steps - 1)],
            "dimensions": list(shape), "vectorization": vectorization, "program": {}}
    prev = "a"
    for s in range(steps):
        name = "b%d" % s
        prog["program"][name] = {
            "computation_string": ("{n} = {c} * ({p}[i-1,j,k] + {p}[i+1,j,k] + {p}[i,j-1,k] + {p}[i,j+1,k] + "
                                   "{p}[i,j,k-1] + {p}[i,j,k+1])").format(n=name, c=coeff, p=prev),
            "boundary_conditions": {prev: dict(boundary)}, "data_type": dtype}
        prev = name
    return prog


def jacobi2d_chain(shape, steps, dtype="float64", boundary=None, data="constant:1.0", vectorization=1):
    boundary = boundary or {"type": "shrink"}
    prog = {"inputs": {"a": {"data": data, "data_type": dtype}}, "outputs": ["b%d" % (steps - 1)],
            "dimensions": list(shape), "vectorization": vectorization, "program": {}}
    prev = "a"
    for s in range(steps):
        name = "b%d" % s
        prog["program"][name] = {
            "computation_string": "{n} = 0.25 * ({p}[j-1,k] + {p}[j+1,k] + {p}[j,k-1] + {p}[j,k+1])".format(
                n=name, p=prev),
            "boundary_conditions": {prev: dict(boundary)}, "data_type": dtype}
        prev = name
    return prog


def hdiff(shape, dtype="float32"):
    sh = {"type": "shrink"}
    return {
        "inputs": {"inp": {"data": "constant:1.0", "data_type": dtype},
                   "coeff": {"data": "constant:0.025", "data_type": dtype}},
        "outputs": ["out"], "dimensions": list(shape),
        "program": {
            "lap": {"computation_string": "lap = 4.0*inp[i,j,k] - (inp[i+1,j,k] + inp[i-1,j,k] + inp[i,j+1,k] + inp[i,j-1,k])",
                    "boundary_conditions": {"inp": dict(sh)}, "data_type": dtype},
            "flx": {"computation_string": "d = lap[i+1,j,k] - lap[i,j,k]; flx = 0.0 if d*(inp[i+1,j,k] - inp[i,j,k]) > 0.0 else d",
                    "boundary_conditions": {"lap": dict(sh), "inp": dict(sh)}, "data_type": dtype},
            "fly": {"computation_string": "d = lap[i,j+1,k] - lap[i,j,k]; fly = 0.0 if d*(inp[i,j+1,k] - inp[i,j,k]) > 0.0 else d",
                    "boundary_conditions": {"lap": dict(sh), "inp": dict(sh)}, "data_type": dtype},
            "out": {"computation_string": "out = inp[i,j,k] - coeff[i,j,k]*(flx[i,j,k] - flx[i-1,j,k] + fly[i,j,k] - fly[i,j-1,k])",
                    "boundary_conditions": {"inp": dict(sh), "coeff": dict(sh), "flx": dict(sh), "fly": dict(sh)},
                    "data_type": dtype},
        }}


def hdiff_jki(shape, dtype="float32"):
    """The same program in the reference's own COSMO layout convention J,K,I
    (``stencilflow/sdfg_to_stencilflow.py:46-68``): ``shape = [NJ, NK, NI]``, the horizontal plane is
    (iterator i, iterator k), the vertical axis (iterator j) carries no taps (SURVEY section 8d,
    config 3 "run that variant too")."""
    sh = {"type": "shrink"}
    return {
        "inputs": {"inp": {"data": "constant:1.0", "data_type": dtype},
                   "coeff": {"data": "constant:0.025", "data_type": dtype}},
        "outputs": ["out"], "dimensions": list(shape),
        "program": {
            "lap": {"computation_string": "lap = 4.0*inp[i,j,k] - (inp[i,j,k+1] + inp[i,j,k-1] + inp[i+1,j,k] + inp[i-1,j,k])",
                    "boundary_conditions": {"inp": dict(sh)}, "data_type": dtype},
            "flx": {"computation_string": "d = lap[i,j,k+1] - lap[i,j,k]; flx = 0.0 if d*(inp[i,j,k+1] - inp[i,j,k]) > 0.0 else d",
                    "boundary_conditions": {"lap": dict(sh), "inp": dict(sh)}, "data_type": dtype},
            "fly": {"computation_string": "d = lap[i+1,j,k] - lap[i,j,k]; fly = 0.0 if d*(inp[i+1,j,k] - inp[i,j,k]) > 0.0 else d",
                    "boundary_conditions": {"lap": dict(sh), "inp": dict(sh)}, "data_type": dtype},
            "out": {"computation_string": "out = inp[i,j,k] - coeff[i,j,k]*(flx[i,j,k] - flx[i,j,k-1] + fly[i,j,k] - fly[i-1,j,k])",
                    "boundary_conditions": {"inp": dict(sh), "coeff": dict(sh), "flx": dict(sh), "fly": dict(sh)},
                    "data_type": dtype},
        }}


def programs_dir():
    """``programs/`` of the repository: the JSON of the BASELINE.json configurations (tracked)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return os.environ.get("SFB200_PROGRAMS", os.path.join(root, "programs"))


def scratch_dir():
    """Where generated programs that are not one of the tracked configurations go (weak-scaled domains,
    test chains): next to the compiled-kernel cache."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(os.environ.get("SFB200_CACHE", os.path.join(root, ".sfcache")), "programs")
    os.makedirs(d, exist_ok=True)
    return d


def write_program(prog, name, directory=None):
    """Writes ``prog`` as ``<name>.json`` and returns the path: into ``directory`` if given, over the
    tracked file of that name in ``programs/`` if there is one, else into the scratch directory."""
    if directory is None:
        tracked = os.path.join(programs_dir(), name + ".json")
        directory = programs_dir() if os.path.isfile(tracked) else scratch_dir()
    path = os.path.join(directory, name + ".json")
    text = json.dumps(prog, indent=1)
    if not os.path.isfile(path) or open(path).read() != text:
        tmp = path + ".tmp{}".format(os.getpid())
        with open(tmp, "w") as f:
            f.write(text)
        os.replace(tmp, path)
    return path


def jacobi2d_weighted(shape, steps, dtype="float64", boundary=None):
    """BASELINE configs[3], secondary variant: the 2-D chain with every stage
    multiplied by a 1-D weight ``w[k]`` (``input_dims ["k"]``) -- a lower-dimensional input inside a
    fused chain.  The reference front end cannot express this in a 2-D program (its ``"[" -> "[i,"``
    rewrite, ``kernel_chain_graph.py:399-403``); semantics by analogy with the 3-D case."""
    prog = jacobi2d_chain(shape, steps, dtype=dtype, boundary=boundary)
    prog["inputs"]["w"] = {"data": "constant:1.0", "data_type": dtype, "input_dims": ["k"]}
    for op in prog["program"].values():
        op["computation_string"] = op["computation_string"].replace("0.25 *", "0.25 * w[k] *")
    return prog


VARIANTS = {2: ("jki",), 3: ("w1d",)}


def baseline_config(index, variant=None):
    """(name, program dict, halo) for configs[index] of BASELINE.json (``variant``: the secondary
    forms SURVEY section 8d asks to report alongside: "jki" for config 2, "w1d" for config 3)."""
    if variant and variant not in VARIANTS.get(index, ()):
        raise ValueError("config {} has no variant {!r}".format(index, variant))
    if index == 2 and variant == "jki":
        return "hdiff_jki_1024x80x1024_f32", hdiff_jki([1024, 80, 1024]), 2
    if index == 3 and variant == "w1d":
        return "jacobi2d_32768_16itr_f64_shrink_w1d", jacobi2d_weighted([32768, 32768], 16), 16
    if index == 0:
        return "jacobi3d_32x32x32_8itr_8vec", jacobi3d_chain([32, 32, 32], 8, vectorization=4), 0
    if index == 1:
        return "jacobi3d_1024_8itr_f32", jacobi3d_chain([1024, 1024, 1024], 8), 0
    if index == 2:
        return "hdiff_1024x1024x80_f32", hdiff([1024, 1024, 80]), 2
    if index == 3:
        return "jacobi2d_32768_16itr_f64_shrink", jacobi2d_chain([32768, 32768], 16), 16
    if index == 4:
        return "jacobi3d_2048_64itr_f32", jacobi3d_chain([2048, 2048, 2048], 64), 0
    raise IndexError(index)
