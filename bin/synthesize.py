#!/usr/bin/env python3
"""Generate a synthetic stencil program (same arguments as the reference's bin/synthesize.py:34-62).

    synthesize.py data_type num_stages num_fields_spatial size_x size_y size_z extent_x extent_y extent_z
                  [-fork_frequency F] [-fork_length_left N] [-fork_length_right N]
                  [-stencil_shape cross|box|diffusion|hotspot] [-vectorize W] [-o PATH]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from stencilflow_b200 import synthesize as syn  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("data_type", choices=["float32", "float64"])
    ap.add_argument("num_stages", type=int)
    ap.add_argument("num_fields_spatial", type=float,
                    help="fields per stencil read from external memory (fractional numbers allowed)")
    for n in ("size_x", "size_y", "size_z"):
        ap.add_argument(n, type=int, help="domain size (0 drops the dimension)")
    for n in ("extent_x", "extent_y", "extent_z"):
        ap.add_argument(n, type=int, help="stencil extent")
    ap.add_argument("-fork_frequency", type=float, default=0.0, help="rate at which forks are generated")
    ap.add_argument("-fork_length_left", type=int, default=2)
    ap.add_argument("-fork_length_right", type=int, default=2)
    ap.add_argument("-stencil_shape", choices=list(syn.SHAPES), default="cross")
    ap.add_argument("-vectorize", type=int, default=1)
    ap.add_argument("-o", dest="output", default=None, help="output path (default: name derived from the arguments)")
    a = ap.parse_args(argv)
    values = [a.data_type, a.num_stages, a.num_fields_spatial, a.size_x, a.size_y, a.size_z, a.extent_x,
              a.extent_y, a.extent_z, a.fork_frequency, a.fork_length_left, a.fork_length_right,
              a.stencil_shape, a.vectorize]
    program = syn.synthesize(*values)
    path = a.output or syn.output_file_name(*values)
    syn.write(program, path)
    print("Wrote synthetic stencil to: {}".format(path))
    return 0


if __name__ == "__main__":
    sys.exit(main())
