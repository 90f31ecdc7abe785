"""Prints the full-size whole-region parity numbers (tests/test_hotpath_gpu.py::full_region_report): Spot, 26 bones,
256^2 on the GPU against the reference's own LASR.forward lines (tests/golden/region_full_ref.npz)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_hotpath_gpu as T
for fused, inject in ((True, True), (True, False), (False, False)):
    try:
        rep = T.full_region_report(fused, inject)
    except Exception as e:
        rep = {"error": repr(e)[:500]}
    print(json.dumps({"fused": fused, "identical_geometry": inject, **rep}))
