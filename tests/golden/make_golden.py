#!/usr/bin/env python3
"""Freezes known answers for the programs of the reference's test-suite into known_answers.json.

The reference stores no output vectors (its program tests compare FPGA emulation with the CPU SDFG of
the same JSON, ``test/test_stencilflow.py:188-224``), and it cannot be executed in this container
(DaCe 0.10.8 needs Python < 3.10).  The values frozen here were produced by ``oracle/reference_numpy.py``
and cross-checked three ways: (a) the small ``simulator*`` results were derived by hand from the JSON
(see HAND_DERIVED below: these literals are written out here, not computed); (b) they agree with the
independently written C++/OpenMP restatement ``oracle/reference_cpp.py``; (c) they agree with the values
SURVEY.md section 8c lists from a separate throw-away evaluator.
Run from the repository root:  python tests/golden/make_golden.py
"""
import glob
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import reference_numpy as rn  # noqa: E402

# Worked by hand from the program text (boundary constant in brackets):
#  simulator:   res[j,k] = arrA[j,k] + arrA[j,k+2] + 3.14, arrA = [[0,1,2],[3,4,5]], OOB -> 0
#               row0: 0+2+3.14, 1+0+3.14, 2+0+3.14 ; row1: 3+5+3.14, 4+0+3.14, 5+0+3.14
#  simulator2:  5-point sum of ones with OOB -> 0: corners 3, edges 4, centre 5
#  simulator9:  kernelA = arrA+1, kernelB = kernelA+1, res = kernelA+kernelB = 2*arrA+3
#  simulator10: kA=kB=kC=arrA, kD=arrA+1, res = 5*arrA+1
#  simulator11: kA = arrA+1; kB = kA[j,k-100] (always OOB -> 0) + kA[j,k+2]; res = kA + kB
#               arrA=[[0,1,2],[3,4,5],[6,7,8]] -> kA=[[1,2,3],[4,5,6],[7,8,9]]; kB[:,0]=kA[:,2], else 0
HAND_DERIVED = {
    "ref_simulator": {"res": [[5.14, 4.14, 5.14], [11.14, 7.14, 8.14]]},
    "ref_simulator2": {"res": [[3, 4, 3], [4, 5, 4], [3, 4, 3]]},
    "ref_simulator9": {"res": [[3, 5, 7], [9, 11, 13]]},
    "ref_simulator10": {"res": [[1, 6, 11], [16, 21, 26]]},
    "ref_simulator11": {"res": [[4, 2, 3], [10, 5, 6], [16, 8, 9]]},
}


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "programs", "ref_*.json"))):
        name = os.path.splitext(os.path.basename(path))[0]
        res = rn.run_reference(path)
        entry = {}
        for k, v in res.items():
            rec = {"dtype": v.dtype.name, "shape": list(v.shape), "sum": float(v.sum(dtype=np.float64))}
            if v.size <= 64:
                rec["values"] = v.tolist()
            else:
                pts = [tuple(0 for _ in v.shape), tuple(s - 1 for s in v.shape), tuple(s // 2 for s in v.shape),
                       tuple(min(1, s - 1) for s in v.shape), tuple([0] * (v.ndim - 1) + [1])]
                rec["points"] = [[list(p), float(v[p])] for p in pts]
            entry[k] = rec
        out[name] = entry
    for name, fields in HAND_DERIVED.items():
        for k, vals in fields.items():
            got = np.array(out[name][k]["values"])
            assert np.allclose(got, np.array(vals, dtype=float), rtol=1e-14, atol=0), (name, k)
            out[name][k]["hand_derived"] = True
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", len(out), "programs")


if __name__ == "__main__":
    main()
