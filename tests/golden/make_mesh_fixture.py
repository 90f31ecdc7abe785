"""Golden INPUT fixture: the Spot mesh that SURVEY.md section 8(c) names as the reference's only
hot-path-relevant asset.  Reads the OBJ where it lies under /root/reference (this container
only) and stores the raw geometry (no texture) as tests/golden/spot_mesh.npz.

    python tests/golden/make_mesh_fixture.py
"""
import os

import numpy as np

SRC = "/root/reference/database/misc/spot/spot_triangulated.obj"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "spot_mesh.npz")


def main():
    v, f = [], []
    with open(SRC) as fh:
        for line in fh:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                v.append([float(x) for x in p[1:4]])
            elif p[0] == "f":
                f.append([int(tok.split("/")[0]) - 1 for tok in p[1:4]])
    v = np.asarray(v, dtype=np.float32)
    f = np.asarray(f, dtype=np.int32)
    assert v.shape == (2930, 3) and f.shape == (5856, 3), (v.shape, f.shape)
    np.savez_compressed(DST, vertices=v, faces=f, source=np.array(SRC))
    print("wrote", DST, v.shape, f.shape)


if __name__ == "__main__":
    main()
