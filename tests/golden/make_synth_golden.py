#!/usr/bin/env python3
"""Golden programs for the synthetic-program generator, produced by the REFERENCE's own script
(``/root/reference/bin/synthesize.py`` runs in the build container: it needs only click and numpy).
The files under tests/golden/synth/ are its unmodified outputs; tests/test_synthesize.py requires
``stencilflow_b200.synthesize`` to reproduce them byte for byte.  CASES is shared with that test.
Run from the repository root (build container only):  python tests/golden/make_synth_golden.py
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "synth")
REF_SCRIPT = "/root/reference/bin/synthesize.py"

CASES = [
    "float32 8 0 32 32 32 1 1 1",
    "float64 16 0 64 64 0 1 1 0",
    "float32 4 0.5 16 16 16 1 1 1 -stencil_shape box",
    "float32 5 1 12 12 12 2 1 1 -stencil_shape diffusion",
    "float32 6 0 16 16 16 1 1 1 -stencil_shape hotspot",
    "float64 4 0.5 32 32 0 1 1 0 -stencil_shape hotspot",
    "float32 6 0.3 16 16 16 1 1 1 -fork_frequency 0.5 -fork_length_left 1 -fork_length_right 3",
    "float32 5 0 24 16 16 1 2 0 -stencil_shape box -fork_frequency 0.34 -vectorize 4",
    "float64 3 2 10 10 10 1 0 1 -stencil_shape diffusion -fork_frequency 1",
    "float32 4 0 256 0 0 2 0 0",
    "float32 3 0 0 48 64 1 1 0",
    "float64 2 1.5 8 8 8 0 0 2 -stencil_shape box",
]

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for case in CASES:
        subprocess.run([sys.executable, REF_SCRIPT] + case.split(), cwd=OUT, check=True)
