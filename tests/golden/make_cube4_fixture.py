"""Parses the reference's only data file, data/cube4.dat (an ALBERT macro triangulation: 125 vertices on the 5^3
lattice of [0,1]^3, 384 tetrahedra; reference data/cube4.dat:1-6, sections `vertex coordinates` / `element vertices`)
into tests/golden/cube4_mesh.json -- the input of BASELINE.json's config 1 (SURVEY.md 8d "Config 1 restated").
Run in the build container (the reference tree does not travel):

    python tests/golden/make_cube4_fixture.py

Vertex coordinates are stored as integers in units of 1/4 (exact), elements as 4 vertex indices."""
import json
import re
from pathlib import Path

import numpy as np

SRC = Path("/root/reference/data/cube4.dat")


def main():
    txt = SRC.read_text()
    nv = int(re.search(r"number of vertices:\s*(\d+)", txt).group(1))
    ne = int(re.search(r"number of elements:\s*(\d+)", txt).group(1))

    def section(name, nxt):
        a = txt.index(name) + len(name)
        return txt[a:txt.index(nxt, a)]

    V = np.array(section("vertex coordinates:", "element vertices:").split(), float).reshape(nv, 3)
    T = np.array(section("element vertices:", "element boundaries:").split(), int).reshape(ne, 4)
    q = np.rint(V * 4).astype(int)
    assert np.array_equal(q / 4.0, V) and T.min() == 0 and T.max() == nv - 1
    out = {"source": "reference data/cube4.dat (ALBERT macro triangulation)", "unit": 0.25,
           "vertices_quarter_units": q.tolist(), "elements": T.tolist()}
    (Path(__file__).parent / "cube4_mesh.json").write_text(json.dumps(out, separators=(",", ":")))
    print(nv, "vertices", ne, "elements")


if __name__ == "__main__":
    main()
