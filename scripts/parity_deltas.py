#!/usr/bin/env python
"""Tier-B (device GCG) against every recorded reference run of tests/golden/gcg_reference.json:
prints, per case, the reference's and the device solver's outer-iteration count, their difference
and the largest relative eigenvalue difference.  The north_star contract is |delta| <= 1
(BASELINE.md section 4); the log of this script on B200 is kept under profiles/.

    python scripts/parity_deltas.py [> profiles/parity_deltas_rNN.log]
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from gcge_b200 import api, problems as P  # noqa: E402

ARGV_TO_PARAM = {"-gcge_compW_cg_order": ("compW_cg_order", int),
                 "-gcge_compW_cg_auto_shift": ("compW_cg_auto_shift", int),
                 "-gcge_compW_cg_shift": ("compW_cg_shift", float),
                 "-gcge_initX_orth_method": ("initX_orth_method", str),
                 "-gcge_compP_orth_method": ("compP_orth_method", str),
                 "-gcge_compW_orth_method": ("compW_orth_method", str)}
ORTH = {"mgs": 0, "bgs": 1}


def overrides(argv):
    out = {}
    for i in range(0, len(argv), 2):
        name, conv = ARGV_TO_PARAM[argv[i]]
        v = conv(argv[i + 1])
        out[name] = ORTH[v] if conv is str else v
    return out


def main():
    api.init(0)
    gold = json.loads((ROOT / "tests" / "golden" / "gcg_reference.json").read_text())["cases"]
    worst = 0
    print(f"{'case':>4} {'generator':>18} {'args':>12} {'nev':>4} {'ref_it':>6} {'dev_it':>6} {'delta':>5} "
          f"{'ref_conv':>8} {'dev_conv':>8} {'max_rel_eval':>12}  options")
    for i, c in enumerate(gold):
        pen = getattr(P, c["generator"])(**c["args"])
        A = api.Mat(pen.A); B = None if pen.B is None else api.Mat(pen.B)
        try:
            o = api.gcg_solve(A, B, nev=c["nev"], **overrides(c.get("argv", [])))
        except (AttributeError, api.B200Error) as e:
            print(f"{i:>4} {c['generator']:>18} {json.dumps(c['args']):>12} {c['nev']:>4}  not run: {e}")
            continue
        k = min(o["nev_conv"], c["nev_conv"])
        ref = np.array(c["eval"][:k])
        err = float(np.max(np.abs(o["eval"][:k] - ref) / np.abs(ref)))
        d = o["num_iter"] - c["num_iter"]
        worst = max(worst, abs(d))
        print(f"{i:>4} {c['generator']:>18} {json.dumps(c['args']):>12} {c['nev']:>4} {c['num_iter']:>6} "
              f"{o['num_iter']:>6} {d:>+5} {c['nev_conv']:>8} {o['nev_conv']:>8} {err:>12.2e}  {' '.join(c.get('argv', []))}")
        A.close()
        if B is not None:
            B.close()
    print(f"worst |delta| = {worst}")


if __name__ == "__main__":
    main()
