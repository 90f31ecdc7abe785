"""Host build of the CUDA kernels' per-(pixel, face) arithmetic (tests/host_harness): the predicted-edge shortcut
for pixels inside a face (raster_common.cuh: euclid_distance, LASR_FAST_INSIDE) must select the very edge the
reference's three-edge loop selects -- every bit of (w, t, dx, dy, dis, D) equal -- on real meshes and on
pathological triangle soups, with the fallback to the full loop staying rare on real meshes.  No GPU needed."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402

DIST_EPS = float(np.float32(np.log(1 / 1e-4 - 1)))


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    d = tmp_path_factory.mktemp("host_harness")
    out = {}
    for v in (0, 1):
        so = str(d / f"raster_math_host_{v}.so")
        subprocess.check_call([os.path.join(ROOT, "tests", "host_harness", "build.sh"), so, f"-DLASR_FAST_INSIDE={v}"])
        out[v] = ctypes.CDLL(so)
        assert out[v].lasr_host_fast_inside() == v
    return out


def _run(L, f9, px, py, sigma):
    n = len(px)
    out = np.zeros((n, 10), np.float32); fr = np.zeros(n, np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L.lasr_host_soft_fragment(P(f9), P(px), P(py), ctypes.c_int(n), ctypes.c_float(sigma), ctypes.c_float(DIST_EPS), P(out), P(fr))
    return out, fr


def _pairs(fv, S, sigma, rng, cap=400000):
    """All grid pixels inside every face's dilated bounding box (sub-sampled for huge faces)."""
    r = np.sqrt(DIST_EPS * sigma)
    coords = ((2 * np.arange(S) + 1 - S) / S).astype(np.float32)
    F, X, Y = [], [], []
    for f in fv:
        xs, ys = f[[0, 3, 6]], f[[1, 4, 7]]
        cx = coords[(coords >= xs.min() - r) & (coords <= xs.max() + r)]
        cy = coords[(coords >= ys.min() - r) & (coords <= ys.max() + r)]
        if len(cx) == 0 or len(cy) == 0:
            continue
        if len(cx) * len(cy) > 2500:
            cx, cy = rng.choice(cx, 50), rng.choice(cy, 50)
        gx, gy = np.meshgrid(cx, cy)
        F.append(np.repeat(f[None], gx.size, 0)); X.append(gx.ravel()); Y.append(gy.ravel())
    F = np.concatenate(F).astype(np.float32); X = np.concatenate(X).astype(np.float32); Y = np.concatenate(Y).astype(np.float32)
    if len(X) > cap:
        s = rng.choice(len(X), cap, replace=False)
        F, X, Y = F[s], X[s], Y[s]
    return np.ascontiguousarray(F), np.ascontiguousarray(X), np.ascontiguousarray(Y)


def _soup(rng, n, kind):
    p = rng.uniform(-0.8, 0.8, (n, 3, 2)).astype(np.float32)
    if kind == "sliver":
        p[:, 2] = p[:, 0] + (p[:, 1] - p[:, 0]) * rng.uniform(0, 1, (n, 1)).astype(np.float32) + rng.normal(0, 1e-5, (n, 2)).astype(np.float32)
    if kind == "needle":
        p[:, 1] = p[:, 0] + rng.normal(0, 1e-4, (n, 2)).astype(np.float32)
    if kind == "tiny":
        p[:, 1:] = p[:, :1] + rng.normal(0, 2e-3, (n, 2, 2)).astype(np.float32)
    if kind == "degenerate":
        p[:, 2] = p[:, 1]
    if kind == "iso":      # grid-aligned isosceles triangles: exact ties at pixel centres
        S = 64
        c = ((2 * rng.randint(8, 56, (n, 2)) + 1 - S) / S).astype(np.float32)
        h = (rng.randint(1, 8, (n,)) * 2 / S).astype(np.float32)
        p[:, 0] = c + np.stack([-h, -h], 1); p[:, 1] = c + np.stack([h, -h], 1); p[:, 2] = c + np.stack([0 * h, h], 1)
    z = rng.uniform(5, 15, (n, 3, 1)).astype(np.float32)
    return np.concatenate([p, z], 2).reshape(n, 9)


def _check(libs, F, X, Y, sigma, tag):
    a, fa = _run(libs[0], F, X, Y, sigma)
    libs[1].lasr_host_stats(ctypes.byref(ctypes.c_longlong()), ctypes.byref(ctypes.c_longlong()), 1)
    b, fb = _run(libs[1], F, X, Y, sigma)
    inside, slow = ctypes.c_longlong(), ctypes.c_longlong()
    libs[1].lasr_host_stats(ctypes.byref(inside), ctypes.byref(slow), 1)
    same = (a.view(np.uint32) == b.view(np.uint32)).all(1) & (fa.view(np.uint32) == fb.view(np.uint32))
    assert same.all(), f"{tag}: {int((~same).sum())} of {len(X)} pairs differ"
    return inside.value, slow.value


@pytest.mark.parametrize("mesh,S", [("spot", 256), ("ico2", 64), ("bumpy2", 96)])
@pytest.mark.parametrize("sigma", [1e-4, 1e-5])
def test_predicted_edge_is_bit_identical_on_meshes(libs, mesh, S, sigma):
    c = cases.lasr_case(mesh, 2, S, sigma=sigma)
    F, X, Y = _pairs(c["face_vertices"].reshape(-1, 9), S, sigma, np.random.RandomState(0))
    inside, slow = _check(libs, F, X, Y, sigma, mesh)
    assert inside > 1000 and slow < 0.15 * inside      # the shortcut is taken for the vast majority of inside pixels


@pytest.mark.parametrize("kind", ["random", "sliver", "needle", "tiny", "degenerate", "iso"])
def test_predicted_edge_is_bit_identical_on_pathological_soups(libs, kind):
    rng = np.random.RandomState(3)
    F, X, Y = _pairs(_soup(rng, 1500, kind), 64, 1e-4, rng)
    _check(libs, F, X, Y, 1e-4, kind)


# ---- the forward kernel's own pair code on the host -------------------------------------------------------------------
@pytest.mark.parametrize("mesh,S,sigma,bg", [("ico2", 64, 1e-4, (0, 0, 0)), ("bumpy2", 96, 1e-4, (0, 0, 0)),
                                             ("ico2", 64, 1e-5, (1, 1, 1)), ("spot", 128, 1e-4, (0, 0, 0))])
def test_forward_pair_code_run_on_the_host_reproduces_the_oracle(libs, mesh, S, sigma, bg):
    """forward_pair<kFast, K = 1> -- the function the forward kernel inlines: border test, record load, soft fragment,
    alpha product, clipped-barycentric depth, online depth softmax, colour blend -- compiled for the host and driven
    pixel by pixel over all faces in ascending order (lasr_host_forward) must give the oracle's image.  A CPU-side
    regression net for edits to the kernel arithmetic; the GPU tests check the same code as compiled by nvcc."""
    from oracle import oracle as O
    c = cases.lasr_case(mesh, 1, S, sigma=sigma, bg=bg)
    r = O.soft_rasterize_forward(c["face_vertices"], c["textures"], c["settings"], np.float32)
    fv = np.ascontiguousarray(c["face_vertices"][0].reshape(-1, 9), np.float32)
    tx = np.ascontiguousarray(c["textures"][0].reshape(-1, 9), np.float32)
    s = c["settings"]
    out = np.zeros((4, S, S), np.float32); agg = np.zeros((2, S, S), np.float32)
    bgv = np.asarray(bg, np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    F32 = ctypes.c_float
    libs[1].lasr_host_forward(P(fv), P(tx), ctypes.c_int(len(fv)), ctypes.c_int(S), F32(s["near"]), F32(s["far"]), F32(sigma),
                              F32(s["gamma_val"]), F32(s["eps"]), F32(DIST_EPS), P(bgv), P(out), P(agg))
    d = np.abs(out - r["soft_colors"][0])
    assert d.max() < 5e-5 and int((d > 1e-4 * np.maximum(1, np.abs(r["soft_colors"][0]))).sum()) == 0
    np.testing.assert_allclose(agg[1], r["aggrs_info"][0][1], rtol=1e-5, atol=1e-7)   # running depth maximum
    np.testing.assert_allclose(agg[0], r["aggrs_info"][0][0], rtol=5e-4)        # softmax sums


# ---- the backward kernel's own pixel-loop body on the host ------------------------------------------------------------
@pytest.mark.parametrize("mesh,S,sigma", [("ico2", 64, 1e-4), ("bumpy2", 96, 1e-4), ("ico2", 64, 1e-5), ("spot", 128, 1e-4)])
def test_backward_pair_code_run_on_the_host_reproduces_the_oracle_gradients(libs, mesh, S, sigma):
    """raster_backward_pair.inc -- the body of the backward kernel's pixel loop, textually included by the kernel and by
    the host harness -- swept over every face's pixel window on the CPU (lasr_host_backward) must give the oracle's
    vertex and colour gradients (north-star tolerance 1e-3; measured ~3e-5)."""
    from oracle import oracle as O
    import parity
    c = cases.lasr_case(mesh, 1, S, sigma=sigma)
    r = O.soft_rasterize_forward(c["face_vertices"], c["textures"], c["settings"], np.float32)
    gf, gt = O.soft_rasterize_backward(r, c["grad_soft_colors"])
    fv = np.ascontiguousarray(c["face_vertices"][0].reshape(-1, 9), np.float32)
    tx = np.ascontiguousarray(c["textures"][0].reshape(-1, 9), np.float32)
    s, n = c["settings"], len(fv)
    ogf = np.zeros((n, 9), np.float32); ogt = np.zeros((n, 9), np.float32)
    scol = np.ascontiguousarray(r["soft_colors"][0]); agg = np.ascontiguousarray(r["aggrs_info"][0])
    gsc = np.ascontiguousarray(c["grad_soft_colors"][0])
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    F32 = ctypes.c_float
    libs[1].lasr_host_backward(P(fv), P(tx), ctypes.c_int(n), ctypes.c_int(S), F32(s["near"]), F32(s["far"]), F32(sigma),
                               F32(s["gamma_val"]), F32(s["eps"]), F32(DIST_EPS), P(scol), P(agg), P(gsc), P(ogf), P(ogt))
    for ours, ref in ((ogf, gf.reshape(n, 9)), (ogt, gt.reshape(n, 9))):
        rep = parity.grad_report(ours, ref)
        assert rep["rel_l2"] < 2e-4 and rep["max_over_scale"] < 3e-4 and rep["frac_gt_1e3"] == 0, rep


# ---- the generic instantiation (every operator mode) on the host, against the reference kernels' own outputs ----------
MODE_CASES = [n for n in cases.GOLDEN_CASES if n.startswith("mode_")]
DIST_ID = {"hard": 0, "barycentric": 1, "euclidean": 2}
RGB_ID = {"hard": 0, "softmax": 1}
ALPHA_ID = {"hard": 0, "sum": 1, "prod": 2}
SAMPLE_ID = {"surface": 0, "vertex": 1}


@pytest.mark.parametrize("name", MODE_CASES)
def test_generic_pair_code_run_on_the_host_reproduces_the_reference_kernels(libs, name):
    """forward_pair<kFast = false> and raster_backward_pair.inc with kFast = false -- the code every operator mode other
    than LASR's runs ('hard' / barycentric distances, 'hard' z-buffer colours, 'sum' / 'hard' alpha, 'surface' texels,
    single-sided faces; IEEE divisions and expf) -- driven on the CPU over the golden mode cases must reproduce what the
    REFERENCE's kernels produced for them on a B200 (tests/golden/softras_ref_mode_*.npz): images with no pixel beyond
    1e-4, gradients within 1e-3."""
    import parity
    g = np.load(os.path.join(ROOT, "tests", "golden", f"softras_ref_{name}.npz"))
    c = cases.GOLDEN_CASES[name]()
    s = c["settings"]
    N, F = c["face_vertices"].shape[:2]
    S, T = s["image_size"], c["textures"].shape[2]
    ids = (DIST_ID[s["dist_func"]], RGB_ID[s["aggr_func_rgb"]], ALPHA_ID[s["aggr_func_alpha"]], SAMPLE_ID[s["texture_type"]],
           int(bool(s["fill_back"])))
    de = float(np.float32(np.log(1.0 / s["dist_eps"] - 1.0)))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    F32, I = ctypes.c_float, ctypes.c_int
    scal = (F32(s["near"]), F32(s["far"]), F32(s["sigma_val"]), F32(s["gamma_val"]), F32(s["eps"]), F32(de)) + tuple(I(v) for v in ids)
    bg = np.asarray(s["background_color"], np.float32)
    img = np.zeros((N, 4, S, S), np.float32); agg = np.zeros((N, 2, S, S), np.float32)
    gf = np.zeros((N, F, 9), np.float32); gt = np.zeros((N, F, T, 3), np.float32)
    for n in range(N):
        fv = np.ascontiguousarray(c["face_vertices"][n].reshape(F, 9), np.float32)
        tx = np.ascontiguousarray(c["textures"][n], np.float32)
        libs[1].lasr_host_forward_modes(P(fv), P(tx), I(F), I(S), I(T), *scal, P(bg), P(img[n]), P(agg[n]))
        gsc = np.ascontiguousarray(c["grad_soft_colors"][n])
        libs[1].lasr_host_backward_modes(P(fv), P(tx), I(F), I(S), I(T), *scal, P(img[n]), P(agg[n]), P(gsc), P(gf[n]), P(gt[n]))
    ir = parity.image_report(img, g["soft_colors"], rel=1e-4)
    assert ir["n_bad"] == 0, (name, ir)
    gfr = parity.grad_report(gf.reshape(g["grad_faces"].shape), g["grad_faces"])
    gtr = parity.grad_report(gt.reshape(g["grad_textures"].shape), g["grad_textures"])
    if gfr["scale"] > 1e-20:
        assert gfr["rel_l2"] < 1e-3 and gfr["max_over_scale"] < 1e-3, (name, gfr)
    assert gtr["rel_l2"] < 1e-3 and gtr["max_over_scale"] < 1e-3, (name, gtr)


# ---- several colour sets riding one geometry pass (K = 2, 3), on the host ---------------------------------------------
@pytest.mark.parametrize("K", [1, 2, 3])
def test_colour_set_kernels_on_the_host_equal_separate_passes(libs, K):
    """forward_pair<true, K> / the backward body with K colour sets (and a detached first set, KF = 1) against K separate
    single-set passes of the same code: image planes and colour gradients bit-identical per set, the silhouette and the
    aggregates shared, the vertex gradient the sum over the sets.  The CPU twin of
    tests/test_raster_gpu.py::test_colour_sets_in_one_pass_equal_separate_renders."""
    c = cases.lasr_case("bumpy2", 1, 64)
    s, S = c["settings"], 64
    fv = np.ascontiguousarray(c["face_vertices"][0].reshape(-1, 9), np.float32)
    F = len(fv)
    rng = np.random.RandomState(5)
    tex = rng.rand(K, F, 9).astype(np.float32)
    tex[0] = c["textures"][0].reshape(F, 9)
    bg = rng.rand(3 * K).astype(np.float32)
    gimg = (rng.randn(3 * K + 1, S, S) * 1e-3).astype(np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    F32, I = ctypes.c_float, ctypes.c_int
    scal = (F32(s["near"]), F32(s["far"]), F32(s["sigma_val"]), F32(s["gamma_val"]), F32(s["eps"]), F32(DIST_EPS))
    L = libs[1]
    texK = np.ascontiguousarray(tex.transpose(1, 0, 2).reshape(F, 9 * K))            # [F][9K], set-major per face
    img = np.zeros((3 * K + 1, S, S), np.float32); agg = np.zeros((2, S, S), np.float32)
    assert L.lasr_host_forward_sets(P(fv), P(texK), I(F), I(S), I(K), *scal, P(bg), P(img), P(agg)) == 0
    singles = []
    for k in range(K):
        im1 = np.zeros((4, S, S), np.float32); ag1 = np.zeros((2, S, S), np.float32)
        t1 = np.ascontiguousarray(tex[k]); b1 = np.ascontiguousarray(bg[3 * k:3 * k + 3])
        assert L.lasr_host_forward_sets(P(fv), P(t1), I(F), I(S), I(1), *scal, P(b1), P(im1), P(ag1)) == 0
        assert np.array_equal(im1[:3].view(np.uint32), img[3 * k:3 * k + 3].view(np.uint32)), k
        assert np.array_equal(im1[3].view(np.uint32), img[3 * K].view(np.uint32))
        assert np.array_equal(ag1.view(np.uint32), agg.view(np.uint32))
        singles.append((t1, im1, ag1))
    for KF in ((0,) if K == 1 else (0, 1)):
        gf = np.zeros((F, 9), np.float32); gt = np.zeros((F, 9 * K), np.float32)
        assert L.lasr_host_backward_sets(P(fv), P(texK), I(F), I(S), I(K), I(KF), *scal, P(img), P(agg), P(gimg), P(gf), P(gt)) == 0
        gf_sum = np.zeros((F, 9), np.float64)
        for k, (t1, im1, ag1) in enumerate(singles):
            g1 = np.zeros((4, S, S), np.float32)
            if k >= KF:
                g1[:3] = gimg[3 * k:3 * k + 3]
            if k == K - 1:
                g1[3] = gimg[3 * K]                                   # the silhouette's gradient enters once
            a = np.zeros((F, 9), np.float32); b = np.zeros((F, 9), np.float32)
            assert L.lasr_host_backward_sets(P(fv), P(t1), I(F), I(S), I(1), I(0), *scal, P(im1), P(ag1), P(g1), P(a), P(b)) == 0
            gf_sum += a
            got = gt[:, 9 * k:9 * k + 9]
            if k < KF:
                assert not got.any()                                  # detached set: no colour gradient
            else:
                assert np.array_equal(got.view(np.uint32), b.view(np.uint32)), (KF, k)
        scale = np.abs(gf_sum).max()
        assert scale > 0 and np.abs(gf - gf_sum).max() < 2e-5 * scale, (KF, np.abs(gf - gf_sum).max(), scale)
        # the PACKED branch of the pair code (per-pixel float4s as raster_pack_kernel / head_backward_kernel<packed> write
        # them) in the two instantiations the library launches: same colour gradients bit for bit (they do not depend
        # on the re-associated sum), vertex gradients up to sum_k g_k (c_k - image_k) = sum_k g_k c_k - T
        if (K, KF) in ((1, 0), (3, 1)):
            pf = np.zeros((F, 9), np.float32); pt = np.zeros((F, 9 * K), np.float32)
            assert L.lasr_host_backward_packed(P(fv), P(texK), I(F), I(S), I(K), I(KF), *scal, P(img), P(agg), P(gimg), P(pf), P(pt)) == 0
            assert np.array_equal(pt.view(np.uint32), gt.view(np.uint32)), (K, KF)
            assert np.abs(pf - gf).max() < 2e-5 * np.abs(gf).max(), (K, KF, np.abs(pf - gf).max(), np.abs(gf).max())
            assert L.lasr_host_backward_packed(P(fv), P(texK), I(F), I(S), I(2), I(0), *scal, P(img), P(agg), P(gimg), P(pf), P(pt)) == 1


@pytest.mark.parametrize("S", [7, 64, 256, 513, 2048, 8192])
def test_pixel_windows_contain_every_pixel_the_exact_predicate_accepts(libs, S):
    """coord_window (raster_common.cuh) turns a face's dilated box into the integer pixel window the forward's tile and
    footprint culls and the backward's sweep use; the exact float predicate of the reference (.cu:32-38) is applied per
    pixel afterwards, so the window only has to be a superset -- for every image size up to the supported maximum, for
    boxes on pixel centres, half-open at the image border, far outside, one ulp wide, +-inf (clamped to the screen) and NaN (culled)."""
    rng = np.random.RandomState(S)
    coords = np.zeros(S, np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    libs[1].lasr_host_pixel_coords(ctypes.c_int(S), P(coords))
    assert np.array_equal(coords, ((2.0 * np.arange(S) + 1.0 - S) / S).astype(np.float32))    # .cu:345 in double, rounded once
    lo = rng.uniform(-1.3, 1.3, 4000).astype(np.float32)
    hi = (lo + np.abs(rng.standard_cauchy(4000)).astype(np.float32) * 0.02).astype(np.float32)
    c = rng.choice(coords, 1000)
    lo = np.concatenate([lo, c, c, np.nextafter(c, np.float32(-9)), c - 3e-8, [-5, 2, -np.inf, np.nan, 0.0, -1.0]]).astype(np.float32)
    hi = np.concatenate([hi, c, np.nextafter(c, np.float32(9)), c, c + 3e-8, [-4, 3, 0.0, 1.0, np.inf, 1.0]]).astype(np.float32)
    n = len(lo)
    out = np.zeros((n, 2), np.int32)
    libs[1].lasr_host_coord_windows(P(lo), P(hi), ctypes.c_int(n), ctypes.c_int(S), P(out))
    with np.errstate(invalid="ignore"):
        accept = (coords[None, :] >= lo[:, None]) & (coords[None, :] <= hi[:, None])           # exact predicate per pixel
    idx = np.arange(S)[None, :]
    inside = (idx >= out[:, :1]) & (idx <= out[:, 1:])
    nan = np.isnan(lo) | np.isnan(hi)
    assert not (accept & ~inside)[~nan].any()                                                  # superset, +-inf bounds included
    assert (out[nan, 0] > out[nan, 1]).all()                                                    # NaN boxes: empty window
    assert (out[~nan, 0] >= 0).all() and (out[~nan, 1] <= S - 1).all()
    assert np.array_equal(out[np.isneginf(lo) & (hi == 0.0)][0], [0, np.searchsorted(coords, 0.0, "right") - 1])   # clamped, not culled
    width = (out[:, 1] - out[:, 0] + 1).clip(0)[~nan]
    assert (width <= accept.sum(1)[~nan] + 2).all()                                             # tight: at most one spare pixel per side
    assert (width == accept.sum(1)[~nan])[:4000].mean() > 0.9                                   # ... and usually none


@pytest.mark.parametrize("kind", ["random", "sliver", "needle", "tiny", "degenerate", "iso"])
def test_kernel_code_matches_the_oracle_on_pathological_soups(libs, kind):
    """Whole images and gradients of triangle soups -- slivers, needles, sub-pixel and exactly degenerate faces (a
    zero-length edge: 0/0 in the closest-point parameter, which CUDA's fmax-style clamp turns into 0), grid-aligned
    faces with exact ties -- from the kernels' own forward and backward code on the host against the oracle."""
    from oracle import oracle as O
    import parity
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    F32, I = ctypes.c_float, ctypes.c_int
    for sigma in (1e-4, 1e-5):
        rng = np.random.RandomState(3)
        S, n = 64, 300
        fv = _soup(rng, n, kind).astype(np.float32)
        tx = rng.rand(n, 9).astype(np.float32)
        s = dict(image_size=S, background_color=[0.1, 0.2, 0.3], near=4.0, far=16.0, fill_back=True, eps=1e-3, sigma_val=sigma,
                 dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-2, aggr_func_rgb="softmax", aggr_func_alpha="prod",
                 texture_type="vertex")
        r = O.soft_rasterize_forward(fv.reshape(1, n, 3, 3), tx.reshape(1, n, 3, 3), s, np.float32)
        g = (rng.randn(1, 4, S, S) * 1e-3).astype(np.float32)
        ogf, ogt = O.soft_rasterize_backward(r, g)
        assert np.isfinite(r["soft_colors"]).all() and np.isfinite(ogf).all() and np.isfinite(ogt).all()
        img = np.zeros((4, S, S), np.float32); agg = np.zeros((2, S, S), np.float32)
        bg = np.asarray(s["background_color"], np.float32)
        scal = (F32(s["near"]), F32(s["far"]), F32(sigma), F32(s["gamma_val"]), F32(s["eps"]), F32(DIST_EPS))
        libs[1].lasr_host_forward(P(fv), P(tx), I(n), I(S), *scal, P(bg), P(img), P(agg))
        ir = parity.image_report(img[None], r["soft_colors"], rel=1e-4)
        assert ir["n_bad"] == 0 and ir["max_abs"] < 5e-5, (kind, sigma, ir)
        gf = np.zeros((n, 9), np.float32); gt = np.zeros((n, 9), np.float32)
        sc = np.ascontiguousarray(r["soft_colors"][0]); ag = np.ascontiguousarray(r["aggrs_info"][0])
        libs[1].lasr_host_backward(P(fv), P(tx), I(n), I(S), *scal, P(sc), P(ag), P(np.ascontiguousarray(g[0])), P(gf), P(gt))
        for ours, ref in ((gf, ogf.reshape(n, 9)), (gt, ogt.reshape(n, 9))):
            rep = parity.grad_report(ours, ref)
            assert rep["rel_l2"] < 1e-4 and rep["max_over_scale"] < 2e-4, (kind, sigma, rep)


@pytest.mark.parametrize("kind", ["sliver", "degenerate"])
def test_generic_kernel_code_matches_the_oracle_in_every_mode_combination(libs, kind):
    """All 72 combinations of distance (hard / barycentric / euclidean), colour aggregation (hard / softmax), alpha
    aggregation (hard / sum / prod), texture type (surface / vertex) and fill_back, on pathological soups: the generic
    instantiation of the kernels' forward and backward code (host build) against the oracle."""
    import itertools
    from oracle import oracle as O
    import parity
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    F32, I = ctypes.c_float, ctypes.c_int
    S, n = 32, 70
    for dist, rgb, alpha, tt, fb in itertools.product(DIST_ID, RGB_ID, ALPHA_ID, SAMPLE_ID, (True, False)):
        rng = np.random.RandomState(7)
        fv = _soup(rng, n, kind).astype(np.float32)
        Tt = 3 if tt == "vertex" else 4
        tx = rng.rand(n, Tt, 3).astype(np.float32)
        if tt == "surface":
            fv[-1, [0, 3, 6]] += 10          # the last face's out-of-bounds texel read is undefined in the reference (cases.py)
        sigma = 1e-2 if dist == "barycentric" else 1e-4
        s = dict(image_size=S, background_color=[0.1, 0.2, 0.3], near=4.0, far=16.0, fill_back=fb, eps=1e-3, sigma_val=sigma,
                 dist_func=dist, dist_eps=1e-4, gamma_val=1e-4 if rgb == "hard" else 1e-2, aggr_func_rgb=rgb,
                 aggr_func_alpha=alpha, texture_type=tt)
        r = O.soft_rasterize_forward(fv.reshape(1, n, 3, 3), tx.reshape(1, n, Tt, 3), s, np.float32)
        g = (rng.randn(1, 4, S, S) * 1e-3).astype(np.float32)
        ogf, ogt = O.soft_rasterize_backward(r, g)
        scal = (F32(s["near"]), F32(s["far"]), F32(sigma), F32(s["gamma_val"]), F32(s["eps"]), F32(DIST_EPS)) + \
            tuple(I(v) for v in (DIST_ID[dist], RGB_ID[rgb], ALPHA_ID[alpha], SAMPLE_ID[tt], int(fb)))
        img = np.zeros((4, S, S), np.float32); agg = np.zeros((2, S, S), np.float32)
        bg = np.asarray(s["background_color"], np.float32)
        libs[1].lasr_host_forward_modes(P(fv), P(tx), I(n), I(S), I(Tt), *scal, P(bg), P(img), P(agg))
        gf = np.zeros((n, 9), np.float32); gt = np.zeros((n, Tt, 3), np.float32)
        sc = np.ascontiguousarray(r["soft_colors"][0]); ag = np.ascontiguousarray(r["aggrs_info"][0])
        libs[1].lasr_host_backward_modes(P(fv), P(tx), I(n), I(S), I(Tt), *scal, P(sc), P(ag), P(np.ascontiguousarray(g[0])), P(gf), P(gt))
        tag = (kind, dist, rgb, alpha, tt, fb)
        assert np.isfinite(img).all() and np.isfinite(gf).all() and np.isfinite(gt).all(), tag
        ir = parity.image_report(img[None], r["soft_colors"], rel=1e-4)
        assert ir["n_bad"] == 0, (tag, ir)
        for ours, ref in ((gf, ogf.reshape(n, 9)), (gt, ogt.reshape(n, Tt, 3))):
            rep = parity.grad_report(ours, ref)
            assert rep["scale"] < 1e-20 or (rep["rel_l2"] < 1e-3 and rep["max_over_scale"] < 1e-3), (tag, rep)


def test_kernel_code_matches_the_oracle_across_depth_ranges_sizes_and_scales(libs):
    """Random soups with vertices behind the camera and at z = 0 (NaN in the reference's arithmetic: the NaN pixels must
    be the same ones), far off screen, odd image sizes, very sharp and very wide sigma, empty and inverted depth
    windows: the kernels' own forward / backward code (host build) against the oracle."""
    from oracle import oracle as O
    import parity
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    F32, I = ctypes.c_float, ctypes.c_int
    rng0 = np.random.RandomState(5)
    base = _soup(rng0, 300, "random").astype(np.float32)
    zcols, xycols = [2, 5, 8], [0, 1, 3, 4, 6, 7]

    def variant(tag):
        f = base.copy()
        if tag == "behind":
            f[:, zcols] = rng0.uniform(-2, 20, (300, 3))
        elif tag == "z0":
            f[:, zcols] = rng0.uniform(0.5, 3, (300, 3)); f[::7, 2] = 0.0
        elif tag == "off4":
            f[:, xycols] *= 4
        elif tag == "off50":
            f[:, xycols] *= 50
        elif tag == "flat":
            f[:, zcols] = 10.0
        return f
    runs = [("behind", 64, 1.0, 15.0, 1e-4), ("z0", 64, 0.1, 5.0, 1e-4), ("off4", 64, 4.0, 16.0, 1e-4),
            ("off50", 64, 4.0, 16.0, 1e-4), ("base", 37, 4.0, 16.0, 1e-4), ("base", 130, 4.0, 16.0, 1e-6),
            ("base60", 48, 4.0, 16.0, 1e-2), ("flat", 64, 4.0, 16.0, 1e-4), ("base", 64, 16.0, 4.0, 1e-4),
            ("base", 64, 10.0, 10.0, 1e-4)]
    for tag, S, near, far, sigma in runs:
        fv = variant(tag)[:60] if tag == "base60" else variant(tag)
        n = len(fv)
        rng = np.random.RandomState(11)
        tx = rng.rand(n, 9).astype(np.float32)
        s = dict(image_size=S, background_color=[0.1, 0.2, 0.3], near=near, far=far, fill_back=True, eps=1e-3, sigma_val=sigma,
                 dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-2, aggr_func_rgb="softmax", aggr_func_alpha="prod",
                 texture_type="vertex")
        r = O.soft_rasterize_forward(fv.reshape(1, n, 3, 3), tx.reshape(1, n, 3, 3), s, np.float32)
        g = (rng.randn(1, 4, S, S) * 1e-3).astype(np.float32)
        ogf, ogt = O.soft_rasterize_backward(r, g)
        img = np.zeros((4, S, S), np.float32); agg = np.zeros((2, S, S), np.float32)
        bg = np.asarray(s["background_color"], np.float32)
        scal = (F32(near), F32(far), F32(sigma), F32(s["gamma_val"]), F32(s["eps"]), F32(DIST_EPS))
        libs[1].lasr_host_forward(P(fv), P(tx), I(n), I(S), *scal, P(bg), P(img), P(agg))
        gf = np.zeros((n, 9), np.float32); gt = np.zeros((n, 9), np.float32)
        sc = np.ascontiguousarray(r["soft_colors"][0]); ag = np.ascontiguousarray(r["aggrs_info"][0])
        libs[1].lasr_host_backward(P(fv), P(tx), I(n), I(S), *scal, P(sc), P(ag), P(np.ascontiguousarray(g[0])), P(gf), P(gt))
        what = (tag, S, near, far, sigma)
        ref_img, rgf, rgt = r["soft_colors"][0], ogf.reshape(n, 9), ogt.reshape(n, 9)
        assert np.array_equal(np.isnan(img), np.isnan(ref_img)), what
        assert np.array_equal(np.isnan(gf), np.isnan(rgf)) and np.array_equal(np.isnan(gt), np.isnan(rgt)), what
        assert (tag == "z0") == bool(np.isnan(ref_img).any()), what
        ok = ~np.isnan(ref_img)
        d = np.abs(img - ref_img)[ok]
        assert d.max() < 5e-5, (what, d.max())
        for ours, ref in ((gf, rgf), (gt, rgt)):
            m = ~np.isnan(ref)
            rep = parity.grad_report(np.where(m, ours, 0), np.where(m, ref, 0))
            assert rep["scale"] < 1e-20 or (rep["rel_l2"] < 1e-4 and rep["max_over_scale"] < 2e-4), (what, rep)


@pytest.mark.parametrize("kind", ["random", "sliver", "needle", "tiny", "degenerate", "iso"])
def test_compact_record_gives_the_backward_the_forwards_own_face(libs, kind):
    """Compact mode keeps 80 bytes per face -- the division results of the record -- and the backward re-derives the
    edge rows and the dilated box from the op's inputs (face_from_compact, raster_common.cuh): every field of the face
    must have the bits the forward's record had, on meshes and on pathological soups (NaN fields of degenerate faces
    included: compared as bit patterns)."""
    rng = np.random.RandomState(len(kind))
    fv = np.ascontiguousarray(_soup(rng, 20000, kind), np.float32)
    if kind == "random":          # + faces reaching far beyond the screen and huge coordinates
        fv[:2000, [0, 1, 3, 4, 6, 7]] *= np.float32(50.0)
        fv[2000:2100, 0] = np.float32(1e30)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for sigma in (1e-4, 1e-5):
        sthr = np.float32(np.sqrt(np.float32(np.log(1.0 / 1e-4 - 1.0)) * np.float32(sigma)))
        assert libs[1].lasr_host_compact_roundtrip(P(fv), ctypes.c_int(len(fv)), ctypes.c_float(float(sthr))) == 0
