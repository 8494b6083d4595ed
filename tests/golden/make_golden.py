"""Generates tests/golden/gcg_reference.json by running the UNMODIFIED reference
(oracle/_ref/libgcge_ref.so, built from /root/reference by oracle/Makefile) on the synthetic
pencils of gcge_b200.problems.  Run in the build container (the reference cannot travel):

    python tests/golden/make_golden.py

Recorded per case: generator name + arguments (the matrices are regenerated bit-identically
from them), the reference's eigenvalues (17 significant digits), iteration count, number of
converged pairs, and the BLAS the reference was linked with (the reference pins none,
SURVEY.md §8c).  Also records slot-level checksums of reference outputs on seeded inputs.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from gcge_b200 import problems as P  # noqa: E402
from oracle import ref  # noqa: E402

CASES = [
    ("laplace1d_pencil", {"n": 807}, 10),      # the reference's built-in driver case
    ("laplace1d_pencil", {"n": 807}, 30),
    ("laplace3d_7pt", {"m": 16}, 10),
    ("laplace3d_7pt", {"m": 20}, 20),
    ("p1_fem_kuhn", {"m": 12}, 10),
    ("p1_fem_kuhn", {"m": 16}, 30),
    ("q1_27pt", {"m": 16}, 20),
    ("laplace3d_7pt", {"m": 30}, 50),
    ("p1_fem_kuhn", {"m": 20}, 50),
    # the block structure of the headline run (nevMax 400, block_size 40, projected problems up to 480)
    ("p1_fem_kuhn", {"m": 24}, 200),
    # SURVEY 8f options (appended: the tests address the cases above by position)
    ("p1_fem_kuhn", {"m": 12}, 10, ("-gcge_compW_cg_order", 2)),
    ("p1_fem_kuhn", {"m": 12}, 10, ("-gcge_initX_orth_method", "bgs", "-gcge_compP_orth_method", "bgs",
                                    "-gcge_compW_orth_method", "bgs")),
    ("p1_fem_kuhn", {"m": 12}, 10, ("-gcge_compW_cg_auto_shift", 1)),
    ("p1_fem_kuhn", {"m": 12}, 10, ("-gcge_compW_cg_shift", 3.0)),
    # BASELINE config 1 restated: P1 on the reference's own mesh data/cube4.dat after two regular
    # refinements (15^3 unknowns), nev = 10; unknowns in lattice order and in the mesh's own vertex order
    ("cube4_p1", {"refine": 2}, 10),
    ("cube4_p1", {"refine": 2, "order": "mesh"}, 10),
]


def main():
    ref.set_threads(1)
    out = {"blas": "OpenBLAS 0.3.15 (opencv_python_headless.libs/libopenblasp-r0-59ffcd50.3.15.so), 1 thread",
           "reference_threads": 1, "cases": []}
    for case in CASES:
        name, kw, nev = case[:3]
        argv = tuple(case[3]) if len(case) > 3 else ()
        pen = getattr(P, name)(**kw)
        r = ref.gcg_solve(pen.A, pen.B, nev=nev, want_evec=False, argv=argv)
        rec = {"generator": name, "args": kw, "nev": nev, "num_iter": r["num_iter"],
               "nev_conv": r["nev_conv"], "eval": [float(f"{v:.17g}") for v in r["eval"][:r["nev_conv"]]]}
        if argv:
            rec["argv"] = [str(a) for a in argv]
        out["cases"].append(rec)
        print(name, kw, nev, argv, r["num_iter"], r["nev_conv"], r["eval"][:3])
    (Path(__file__).parent / "gcg_reference.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
