"""Generate tests/golden/*.npz from the reference tree (run in the build container only).

  python tests/golden/make_golden.py

Reads the reference's input meshes (/root/reference/input/*.obj -- data fixtures, not
source code) and runs the reference's OWN CPU vertex-normal loop, compiled unmodified
into oracle/_ref/libvn_ref.so (oracle/Makefile), to produce golden outputs.  The GPU box
has no /root/reference, so the tests read only the files written here.

Known-answer constants that the reference's tests hold are recorded in KNOWN below with
their citation; tests/test_oracle.py checks the oracle against them.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from rxmesh_b200 import meshio  # noqa: E402

REF_INPUT = "/root/reference/input"
OUT = os.path.dirname(os.path.abspath(__file__))

MESHES = ["sphere3", "dragon", "cube", "bunnyhead", "plane", "plane_5", "torus", "sphere1",
          "diamond"]

KNOWN = {
    # tests/RXMesh_test/test_boundary.cu:27
    "bunnyhead_boundary_vertices": 98,
    # tests/RXMesh_test/test_for_each.cu:25-29 (cube.obj element counts)
    "cube_counts": {"V": 8, "E": 18, "F": 12},
    # reference vertex_normal_ref.h run unmodified (fp32)
    "sphere3_abs_sum": 1299.1416,
    "dragon_sum": 486.098569,
    "dragon_abs_sum": 27471.6973,
    "dragon_n0": [1.6687963, 0.813696265, 1.25155962],
    "dragon_nlast": [-0.788115859, -0.583886743, 1.98309875],
    "sphere3_n0": [-1.35289872, -1.35289884, -1.35289884],
    "sphere3_nlast": [0.511231184, 0.978179276, 2.17772222],
    # sizes of the fixtures
    "sphere3_VF": [386, 768],
    "dragon_VF": [10000, 20000],
}


def main():
    for name in MESHES:
        V, F = meshio.import_obj(os.path.join(REF_INPUT, name + ".obj"))
        out = {"V": V, "F": F}
        if name in ("sphere3", "dragon", "bunnyhead", "torus"):
            out["vn_ref"] = O.ref_vertex_normals(F, V)  # the reference's own loop
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, V.shape, F.shape, sorted(out))
    # patchings saved by the reference itself (cereal archives shipped in input/): golden vectors for the
    # ownership / ribbon rules of the patch builder (tests/test_host_build.py)
    import shutil
    for name in ("sphere3_patches", "torus_patches"):
        shutil.copyfile(os.path.join(REF_INPUT, name), os.path.join(OUT, name))
    with open(os.path.join(OUT, "known_answers.json"), "w") as fh:
        json.dump(KNOWN, fh, indent=1)


if __name__ == "__main__":
    main()
