import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PROGRAMS = os.path.join(ROOT, "tests", "programs")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def program_path(name):
    return os.path.join(PROGRAMS, name + ".json")


def all_programs():
    return sorted(os.path.splitext(f)[0] for f in os.listdir(PROGRAMS) if f.endswith(".json"))


@pytest.fixture(scope="session")
def native_lib():
    """libsfb200.so, built on demand (nvcc cross-compiles without a GPU)."""
    from stencilflow_b200 import build
    build.build_native()
    from stencilflow_b200 import runtime
    return runtime.load_library()


# programs with "shrink" boundaries: width of the border whose content is junk by definition
# (run_program's -halo, reference stencilflow/run_program.py:202-209)
HALO = {
    "hdiff_24x28x16": 2,
    "hdiff_16x20x8_f64": 2,
    "jacobi2d_96x128_6itr_shrink_f64": 6,
    "jacobi3d_24x20x40_4itr_shrink_f64": 4,
    "lowdim3d_20x24x48_3st_shrink_f64": 4,
    "jacobi2d_96x128_6itr_w1d_shrink_f64": 6,
    "sdfgexport_hdiff_jki_48x8x64_f64": 2,
}

# value ranges of the random test inputs; hdiff subtracts a small correction from `inp`, so its
# inputs are kept away from zero to make a *relative* error bound meaningful
INPUT_RANGES = {
    "hdiff_24x28x16": {"inp": (1.0, 2.0), "coeff": (0.0, 0.05)},
    "hdiff_16x20x8_f64": {"inp": (1.0, 2.0), "coeff": (0.0, 0.05)},
    "hdiff_const_10x12x8_f64": {"inp": (1.0, 2.0), "coeff": (0.0, 0.05)},
    "sdfgexport_hdiff_jki_48x8x64_f64": {"inp": (1.0, 2.0), "coeff": (0.0, 0.05), "wgt": (0.9, 1.1)},
    # hotspot: coefficients of a stable explicit step (the generator's default 0.5 is not)
    "synth_hotspot2d_48x64_4st_f64": {"sdc": (0.05, 0.1), "r_x": (0.5, 1.0), "r_y": (0.5, 1.0), "r_z": (0.5, 1.0),
                                      "amb": (0.5, 1.0)},
    "synth_hotspot3d_12x12x16_4st": {s: (0.05, 0.15) for s in ("cc", "cn", "cs", "cw", "ce", "ca", "cb", "sdc")},
    "synth_diffusion_10x12x16_4st": {"c%d" % n: (0.05, 0.25) for n in range(7)},
    # `- wk[k-3]` in b1: keep every intermediate away from zero, otherwise float32 cancellation turns the
    # (legitimate) FMA-contraction differences between nvcc and numpy into relative errors above 1e-5
    "lowdim3d_20x24x48_3st_f32": {"pij": (1.0, 2.0), "pjk": (0.5, 1.5), "pik": (0.5, 1.0), "wk": (0.0, 0.5)},
    "lowdim3d_20x24x48_3st_shrink_f64": {"pij": (1.0, 2.0), "pjk": (0.5, 1.5), "pik": (0.5, 1.0), "wk": (0.0, 0.5)},
}


def random_inputs(name, seed=7):
    """Seeded random inputs for a test program: {input name: ndarray or numpy scalar}."""
    import numpy as np
    from oracle import reference_numpy as rn
    prog = rn.load_program(program_path(name))
    info = rn.ProgramInfo(prog)
    rng = np.random.default_rng(seed)
    inputs = {}
    for field in info.inputs:
        shape = info.field_shape(field)
        dt = info.field_type(field)
        lo, hi = INPUT_RANGES.get(name, {}).get(field, (0.0, 1.0))
        if len(shape) == 0:
            slo, shi = INPUT_RANGES.get(name, {}).get(field, (0.5, 1.5))
            inputs[field] = dt(rng.uniform(slo, shi))
        else:
            inputs[field] = rng.uniform(lo, hi, size=shape).astype(dt)
    return inputs
