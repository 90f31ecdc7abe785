"""The C-ABI library loads and exports every symbol include/lasr_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from lasr_b200 import _build
    return ctypes.CDLL(_build.build_library())


def declared_symbols():
    with open(os.path.join(ROOT, "include", "lasr_b200.h")) as fh:
        src = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(lasr_b200_\w+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "lasr_b200_raster_forward" in syms and "lasr_b200_raster_backward" in syms


def test_every_declared_symbol_is_exported(lib):
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/lasr_b200.h but not exported"


def test_binding_lists_same_symbols():
    from lasr_b200 import _cabi
    assert sorted(_cabi.EXPORTED_SYMBOLS) == declared_symbols()


def test_no_compute_entry_points_without_gpu(lib):
    lib.lasr_b200_abi_version.restype = ctypes.c_int
    assert lib.lasr_b200_abi_version() >= 1
    lib.lasr_b200_raster_workspace_bytes.restype = ctypes.c_size_t
    assert lib.lasr_b200_raster_workspace_bytes(2, 100, 64) >= 2 * 100 * 8
    assert lib.lasr_b200_raster_workspace_bytes(0, 100, 64) == 0


def test_product_never_imports_oracle():
    """The product path must not route through the oracle / any CPU fallback."""
    pkg = os.path.join(ROOT, "lasr_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(d, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(d, f)
                assert "libsoftras_oracle" not in src and "ref_gpu" not in src, os.path.join(d, f)


def test_cpu_tensor_is_rejected_loudly():
    import torch
    from lasr_b200 import _cabi
    import lasr_b200.soft_renderer as sr
    fv = torch.zeros(1, 4, 3, 3)
    tex = torch.zeros(1, 4, 3, 3)
    with pytest.raises(TypeError):
        sr.functional.soft_rasterize(fv, tex, texture_type="vertex")
    with pytest.raises(RuntimeError):
        sr.functional.soft_rasterize(fv, tex, texture_type="vertex")
    assert issubclass(_cabi.CudaTensorRequired, (TypeError, RuntimeError))


def test_rasterizer_rejects_bad_enums():
    import lasr_b200.soft_renderer as sr
    for kw in (dict(dist_func="l1"), dict(aggr_func_rgb="max"), dict(aggr_func_alpha="mean"),
               dict(texture_type="uv")):
        with pytest.raises(ValueError):
            sr.SoftRasterizer(**kw)


def test_every_compute_entry_point_refuses_null_arguments_with_a_message():
    """Argument validation happens before any CUDA call: with null pointers and zero sizes every entry point returns
    non-zero and leaves a message in lasr_b200_last_error() -- no GPU needed, nothing is launched (the reference
    printf's CUDA errors and carries on, soft_rasterize_cuda_kernel.cu:710-712)."""
    from lasr_b200 import _cabi
    L = _cabi.lib()
    before = _cabi.launch_count()
    checked = 0
    for name in _cabi.EXPORTED_SYMBOLS:
        fn = getattr(L, name)
        if fn.argtypes is None or name.endswith("_bytes"):
            continue
        args = [0.0 if t is ctypes.c_float else 0 if t in (ctypes.c_int, ctypes.c_size_t) else None for t in fn.argtypes]
        assert fn(*args) != 0, name
        msg = L.lasr_b200_last_error().decode()
        assert msg and msg.split(":")[0].split("_")[0] in name, (name, msg)
        checked += 1
    assert checked >= 27 and _cabi.launch_count() == before


def test_a_plain_c_program_binds_the_library(tmp_path, lib):
    """The boundary is a C ABI: include/lasr_b200.h must be valid C99 (no C++ in the signatures) and a C program linked
    against liblasr_b200.so must be able to call it -- here the version query and one refused call (null arguments are
    rejected before the device is touched, so this runs without a GPU)."""
    import shutil
    import subprocess
    from lasr_b200 import _build
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "consumer.c"
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "lasr_b200.h"\n'
                   "int main(void) {\n"
                   "  int v = lasr_b200_abi_version();\n"
                   "  int rc = lasr_b200_chamfer_forward(NULL, NULL, NULL, NULL, NULL, NULL, 1, 4, 4, NULL);\n"
                   '  printf("%d %d %s\\n", v, rc, lasr_b200_last_error());\n'
                   "  return (v >= 4 && rc != 0 && strlen(lasr_b200_last_error()) > 0) ? 0 : 1;\n}\n")
    exe = tmp_path / "consumer"
    libdir = os.path.dirname(_build.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-llasr_b200", "-Wl,-rpath," + libdir])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "chamfer_forward" in r.stdout


def test_module_constructors_refuse_what_the_reference_refuses():
    """ValueError for unknown light / camera modes and for eye setters on a projection camera (lighting.py:41-42,
    transform.py:87-99), the texture-mode assertion of SoftRenderer.set_texture_mode (renderer.py:86), and the explicit
    refusals of what is out of scope here."""
    import lasr_b200.soft_renderer as sr
    with pytest.raises(ValueError):
        sr.Lighting(light_mode="pixel")
    with pytest.raises(ValueError):
        sr.Transform(camera_mode="fisheye")
    with pytest.raises(ValueError):
        sr.SoftRenderer(camera_mode="look_at", light_mode="pixel")
    r = sr.SoftRenderer(camera_mode="look_at")
    with pytest.raises(AssertionError):
        r.set_texture_mode("uv")
    with pytest.raises(ValueError):
        sr.SoftRenderer()                       # camera_mode='projection' (the default) needs a [N,3,4] matrix (transform.py:18-19)
    with pytest.raises(NotImplementedError):
        sr.functional.voxelization(None)
    t = sr.Transform(camera_mode="look_at")
    t.set_eyes([0, 0, -3.0])
    assert t.eyes == [0, 0, -3.0]
    import torch
    with pytest.raises(ValueError):
        sr.functional.look_at(torch.zeros(4, 3), [0, 0, -1])          # vertices must be [N,V,3] (look_at.py:16-17)
    with pytest.raises(AssertionError):
        sr.functional.face_vertices(torch.zeros(2, 4, 3), torch.zeros(3, 4, 3, dtype=torch.int32))   # face_vertices.py:10-13 asserts


def test_keep_workspace_switch_has_three_settings(monkeypatch):
    """functional.soft_rasterize.KEEP_WORKSPACE: True (the workspace, 200 B per face), "compact" (80 B per face),
    False (nothing); anything else is refused when a forward reads it.  The compact buffer is smaller than the
    reference's faces_info [N,F,27] at every shape LASR uses, so a maintainer can pass that tensor as `saved`."""
    import sys
    import lasr_b200.soft_renderer as sr
    from lasr_b200 import _cabi
    mod = sys.modules[sr.functional.SoftRasterizeFunction.__module__]
    for value, mode in ((True, "full"), (False, "none"), ("compact", "compact"), (1, "full"), (0, "none")):
        monkeypatch.setattr(mod, "KEEP_WORKSPACE", value)
        assert mod.keep_mode() == mode
    monkeypatch.setattr(mod, "KEEP_WORKSPACE", "small")
    with pytest.raises(ValueError):
        mod.keep_mode()
    L = _cabi.lib()
    for N, F, S in ((64, 5856, 256), (8, 1280, 256), (1, 1280, 256), (32, 20480, 512), (80, 10000, 512)):
        saved = L.lasr_b200_raster_saved_bytes(N, F, S)
        assert saved <= N * F * 27 * 4 and saved >= N * F * 80
        assert L.lasr_b200_raster_backward_scratch_bytes(N, F, S) < L.lasr_b200_raster_workspace_bytes(N, F, S)
    assert L.lasr_b200_raster_saved_bytes(0, 10, 10) == 0


def test_boolean_masks_are_reinterpreted_not_converted():
    """losses._as_u8: a bool mask reaches the kernels as the same bytes (no conversion kernel per step); other dtypes are
    converted."""
    import torch
    from lasr_b200 import losses
    m = torch.rand(3, 5, 7) > 0.5
    u = losses._as_u8(m)
    assert u.dtype == torch.uint8 and u.data_ptr() == m.data_ptr() and torch.equal(u.bool(), m)
    f = (torch.rand(3, 5) > 0.5).float()
    assert losses._as_u8(f).dtype == torch.uint8 and torch.equal(losses._as_u8(f).bool(), f.bool())
    nc = (torch.rand(4, 6) > 0.5).t()
    assert torch.equal(losses._as_u8(nc).bool(), nc)
