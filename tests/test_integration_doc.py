"""The reference-side binding stubs printed in INTEGRATION.md are executed here (no GPU): each stub is extracted from the
document, run against a recording stand-in for `ctypes.CDLL`, and the call it makes is checked against
include/lasr_b200.h -- parameter count, pointer / int / float / size_t positions, and that the reference's scalars land
in the slots the header gives them."""
import ctypes
import os
import re
import types

import pytest
import torch

from test_static_checks import _header_kinds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOC = open(os.path.join(ROOT, "INTEGRATION.md")).read()


def _block(marker):
    """First ```python block after the line containing `marker`."""
    i = DOC.index(marker)
    m = re.search(r"```python\n(.*?)```", DOC[i:], flags=re.S)
    return m.group(1)


class _FakeLib:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if not name.startswith("lasr_b200_"):
            raise AttributeError(name)
        lib = self

        class Fn:
            restype = ctypes.c_int

            def __call__(self, *a):
                lib.calls.append((name, a))
                if name.endswith("_bytes"):
                    return 4096
                if name == "lasr_b200_last_error":
                    return b""
                return 0
        fn = Fn()
        object.__setattr__(self, name, fn)
        return fn


def _kind(a):
    if a is None or isinstance(a, (ctypes.c_void_p, ctypes.c_char_p)):
        return "ptr"
    if isinstance(a, ctypes.c_float):
        return "float"
    if isinstance(a, ctypes.c_size_t):
        return "size"
    if isinstance(a, (bool, int)):
        return "int"
    raise AssertionError(f"argument {a!r} would be passed with ctypes' default conversion")


def _val(a):
    return a.value if hasattr(a, "value") else a


@pytest.fixture()
def fake(monkeypatch):
    lib = _FakeLib()
    monkeypatch.setattr(ctypes, "CDLL", lambda path: lib)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: types.SimpleNamespace(cuda_stream=0x5EED))
    return lib


def test_pybind_level_stub_calls_the_raster_entry_points_as_declared(fake):
    ns = {}
    exec(_block("## B. pybind level"), ns)
    N, F, T, S = 2, 5, 3, 8
    faces, textures = torch.zeros(N, F, 9), torch.zeros(N, F, T, 3)
    finfo, aggrs, img = torch.zeros(N, F, 27), torch.zeros(N, 2, S, S), torch.zeros(N, 4, S, S)
    # the scalars in the order the reference's SoftRasterizeFunction passes them (functional/soft_rasterize.py:55-61)
    scal = (S, 1.5, 77.0, 1e-3, 1e-4, 2, 9.21, 1e-2, 1, 2, 1, True)
    out = ns["forward_soft_rasterize"](faces, textures, finfo, aggrs, img, *scal)
    assert out[0] is finfo and out[1] is aggrs and out[2] is img
    gf, gt, g = torch.zeros_like(faces), torch.zeros_like(textures), torch.zeros_like(img)
    out = ns["backward_soft_rasterize"](faces, textures, img, finfo, aggrs, gf, gt, g, *scal)
    assert out[0] is gf and out[1] is gt
    kinds = _header_kinds()
    calls = {n: a for n, a in fake.calls if n in ("lasr_b200_raster_forward", "lasr_b200_raster_backward")}
    assert set(calls) == {"lasr_b200_raster_forward", "lasr_b200_raster_backward"}
    for name, a in calls.items():
        assert [_kind(x) for x in a] == kinds[name], name
        ptrs = 4 if name.endswith("forward") else 7
        want_ptrs = [faces, textures, aggrs, img] if ptrs == 4 else [faces, textures, img, aggrs, gf, gt, g]
        assert [_val(x) for x in a[:ptrs]] == [t.data_ptr() for t in want_ptrs], name
        v = [_val(x) for x in a[ptrs + 2:]]       # after workspace pointer and size
        assert v[:4] == [N, F, S, T]
        assert v[4:6] == pytest.approx([1.5, 77.0]) and v[6] is None                       # near, far, near_far_dev
        assert v[7:9] == pytest.approx([1e-3, 1e-4]) and v[9] == 2                         # eps, sigma, dist
        assert v[10:12] == pytest.approx([9.21, 1e-2]) and v[12:16] == [1, 2, 1, 1]        # dist_eps, gamma, rgb, alpha, sample, double_side
        assert v[-1] == 0x5EED
        if name.endswith("backward"):
            assert v[16] == 0                                                               # setup_done
    assert ("lasr_b200_raster_workspace_bytes", (N, F, S)) in [(n, tuple(a)) for n, a in fake.calls]


def test_chamfer_stub_calls_the_chamfer_entry_points_as_declared(fake):
    ns = {}
    exec(_block("keeps the reference's Python wrapper can instead"), ns)
    B, n, m = 1, 6, 4
    x1, x2 = torch.zeros(B, n, 3), torch.zeros(B, m, 3)
    d1, d2 = torch.zeros(B, n), torch.zeros(B, m)
    i1, i2 = torch.zeros(B, n, dtype=torch.int32), torch.zeros(B, m, dtype=torch.int32)
    assert ns["forward"](x1, x2, d1, d2, i1, i2) == 1                     # third_party/chamfer3D/dist_chamfer_3D.py:47
    g1, g2, gd1, gd2 = torch.zeros_like(x1), torch.zeros_like(x2), torch.zeros_like(d1), torch.zeros_like(d2)
    assert ns["backward"](x1, x2, g1, g2, gd1, gd2, i1, i2) == 1          # :62-64 (the reference's argument order)
    kinds = _header_kinds()
    (fn, fa), (bn, ba) = fake.calls
    assert fn == "lasr_b200_chamfer_forward" and bn == "lasr_b200_chamfer_backward"
    assert [_kind(x) for x in fa] == kinds[fn] and [_kind(x) for x in ba] == kinds[bn]
    assert [_val(x) for x in fa[:6]] == [t.data_ptr() for t in (x1, x2, d1, d2, i1, i2)] and list(fa[6:9]) == [B, n, m]
    assert [_val(x) for x in ba[:8]] == [t.data_ptr() for t in (x1, x2, gd1, gd2, i1, i2, g1, g2)] and list(ba[8:11]) == [B, n, m]
