"""Synthetic inputs for the LASR render hot path.

Procedural meshes (icosphere), the Spot fixture loader and the camera sweep that
``scripts/render_syn.py:143-173`` of the reference uses to pose Spot.  Pure numpy; used by
``bench.py``, ``__graft_entry__.smoke()`` and the tests to build *inputs* only.
"""
from __future__ import annotations

import os

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPOT_FIXTURE = os.path.join(_REPO, "tests", "golden", "spot_mesh.npz")


def icosphere(level: int, radius: float = 1.0):
    """Unit icosahedron subdivided ``level`` times (level 3 -> 642 V / 1280 F, 4 -> 2562 / 5120,
    5 -> 10242 / 20480), the mesh family LASR initialises from (``--subdivide``,
    nnutils/mesh_net.py:63).  Returns (vertices float32 [V,3], faces int32 [F,3]).

    Same vertex and face counts as the reference's ``meshzoo.iso_sphere(level)`` (third_party/ext_utils/meshzoo.py:229),
    same level-1 mesh; from level 2 on the vertices sit slightly differently on the sphere -- meshzoo refines the flat
    icosahedron ``level`` times and projects once, this projects the midpoints at every level.  Benchmark shapes B / D
    are defined by their counts; the golden cases hash these very vertices, so the construction stays."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
         (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5),
         (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    verts = [np.asarray(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    faces = [tuple(x) for x in f]
    for _ in range(level):
        cache = {}

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        nxt = []
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nxt += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nxt
    return (np.asarray(verts) * radius).astype(np.float32), np.asarray(faces, dtype=np.int32)


def bumpy_sphere(level: int, radius: float = 0.5, seed: int = 0):
    """Icosphere with smooth low-frequency lobes, so that the silhouette has concavities and
    several depth layers overlap (a convex sphere never self-occludes beyond front/back)."""
    v, f = icosphere(level, 1.0)
    rng = np.random.RandomState(seed)
    dirs = rng.randn(6, 3)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    amp = rng.uniform(0.15, 0.45, size=6)
    r = np.ones(len(v))
    for d, a in zip(dirs, amp):
        r += a * np.exp(4.0 * (v @ d - 1.0))
    return (v * r[:, None] * radius).astype(np.float32), f


def load_spot(path: str = SPOT_FIXTURE):
    """Spot (2930 V / 5856 F), vertices preprocessed as scripts/render_syn.py:72-74
    (y *= -1; y += 0.1; /= 1.2).  The fixture is produced by tests/golden/make_mesh_fixture.py."""
    d = np.load(path)
    v = d["vertices"].astype(np.float32).copy()
    v[:, 1] *= -1
    v[:, 1] += 0.1
    v /= 1.2
    return v, d["faces"].astype(np.int32)


def rodrigues(rvec):
    """Axis-angle -> rotation matrix (what cv2.Rodrigues returns, render_syn.py:148)."""
    rvec = np.asarray(rvec, dtype=np.float64)
    th = np.linalg.norm(rvec)
    if th < 1e-12:
        return np.eye(3)
    k = rvec / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def spot_cameras(n_frames: int, seed: int = 0, focal: float = 10.0, depth: float = 10.0):
    """Per-frame rigid pose of scripts/render_syn.py:143-156: rotation
    Rodrigues([rx_i, 3*1.57 + 6.28*i/N, 0]) with rx_0 = 0, rx_i ~ U[0,1); translation (0,0,depth);
    focal 10; principal point 0.  Returns Rmat [N,3,3] such that ``cam = verts @ Rmat + Tmat``
    (the row-vector convention of nnutils/geom_utils.py:45-71), Tmat [N,1,3], pp [N,2], fl [N,1]."""
    rng = np.random.RandomState(seed)
    R = np.zeros((n_frames, 3, 3), dtype=np.float32)
    for i in range(n_frames):
        rx = rng.rand()
        if i == 0:
            rx = 0.0
        R[i] = rodrigues([rx, 3 * 1.57 + 6.28 * i / n_frames, 0.0]).T  # render_syn.py:162-164
    T = np.zeros((n_frames, 1, 3), dtype=np.float32)
    T[:, 0, 2] = depth
    pp = np.zeros((n_frames, 2), dtype=np.float32)
    fl = np.full((n_frames, 1), focal, dtype=np.float32)
    return R, T, pp, fl


def project_frames(verts, Rmat, Tmat, pp, fl):
    """obj->cam + pinhole of the training path (mesh_net.py:298-302): x,y projected, z kept."""
    cam = np.matmul(verts[None].astype(np.float32), Rmat) + Tmat
    out = cam.copy()
    out[:, :, 0] = pp[:, 0:1] + cam[:, :, 0] * fl / cam[:, :, 2]
    out[:, :, 1] = pp[:, 1:2] + cam[:, :, 1] * fl / cam[:, :, 2]
    return out.astype(np.float32), cam.astype(np.float32)


def vertex_colors(n_verts: int, seed: int = 0):
    """sigmoid(N(0,1)) per vertex (third_party/ext_nnutils/mesh_net.py:96,184)."""
    rng = np.random.RandomState(seed)
    return (1.0 / (1.0 + np.exp(-rng.randn(n_verts, 3)))).astype(np.float32)


def render_inputs(verts, faces, n_frames, seed=0):
    """Posed, projected, y-flipped vertices [N,V,3] exactly as the rasterizer sees them in LASR
    (net camera transform is (x, -y, z)), plus near/far by mesh_net.py:304-311."""
    R, T, pp, fl = spot_cameras(n_frames, seed)
    px, cam = project_frames(verts, R, T, pp, fl)
    px[:, :, 1] *= -1
    z = px[:, :, 2]
    dmin, dmax = np.float32(z.min()), np.float32(z.max())
    near = np.float32(dmin - (dmax - dmin) / 2)
    far = np.float32(dmax + (dmax - dmin) / 2)
    return px, cam, float(near), float(far)


def gather_faces(verts_nv3, faces):
    """[N,V,C] x [F,3] -> [N,F,3,C] (functional/face_vertices.py with a shared topology)."""
    return np.ascontiguousarray(verts_nv3[:, faces.astype(np.int64)])


def hot_path_inputs(n_pairs: int, n_bones: int, image_size: int, device, n_hypo: int = 1, seed: int = 0):
    """Synthetic inputs of the whole render hot path (lasr_b200.hotpath.RenderHotPath) at SURVEY 8(d)'s
    config-E shape: the Spot mesh, ``2*n_pairs`` frames on the render_syn.py camera orbit, ``n_bones``
    bones (1 body + parts with small random articulations, rest pose at sampled vertices), random
    observations.  Returns (args tuple for RenderHotPath.forward, list of differentiable leaves)."""
    import torch
    from .geom_utils import quaternion_to_rotation_matrix
    g = torch.Generator().manual_seed(seed)
    v, f = load_spot()
    V, n2, H, B, S = len(v), 2 * n_pairs, n_hypo, n_bones, image_size
    N = n2 * H
    Rb, Tb, _, _ = spot_cameras(N)
    pred_v = torch.from_numpy(v)[None].repeat(N, 1, 1).to(device)
    tex = torch.from_numpy(vertex_colors(V))[None].repeat(N, 1, 1).to(device)
    faces = torch.from_numpy(f)[None].repeat(n2, 1, 1).to(device)
    q = torch.randn(N * B, 4, generator=g) * 0.05
    q[:, 3] = 1
    R = quaternion_to_rotation_matrix(q).view(N, B, 3, 3)
    R[:, 0] = torch.from_numpy(Rb)
    Rmat = R.view(-1, 3, 3).contiguous().to(device)
    T = torch.randn(N * B, 3, generator=g) * 0.01
    T.view(N, B, 3)[:, 0] = torch.from_numpy(Tb[:, 0])
    Tmat = T.to(device)
    ppoint = torch.zeros(n2, 2, device=device)
    scale = torch.full((N, 1), 10.0, device=device)
    bones = None
    if B > 1:
        idx = torch.randperm(V, generator=g)[:H * (B - 1)]
        ctl = torch.from_numpy(v)[idx].to(device)
        bones = dict(ctl_ts=ctl.clone(), rest_ts=ctl.clone(),
                     ctl_rs=torch.tensor([[0., 0, 0, 1]], device=device).repeat(H * (B - 1), 1),
                     log_ctl=torch.ones(H * (B - 1), 3, device=device))
    batch = dict(masks=(torch.rand(n2, S, S, generator=g) > 0.5).float().to(device),
                 imgs=torch.rand(n2, 3, S, S, generator=g).to(device),
                 flow=(torch.randn(n2, 3, S, S, generator=g) * 0.02).to(device),
                 occ=torch.randn(n2, S, S, generator=g).to(device))
    leaves = [pred_v, tex, Rmat, Tmat, ppoint, scale] + (list(bones.values()) if bones else [])
    return dict(pred_v=pred_v, tex=tex, faces=faces, Rmat=Rmat, Tmat=Tmat, ppoint=ppoint, scale=scale, batch=batch,
                bones=bones), leaves


def refine_to_faces(verts, faces, target_faces: int):
    """Refine a closed triangle mesh to exactly ``target_faces`` faces by splitting interior edges at their
    midpoints, longest first, never two edges of the same triangle (each split turns the edge's two triangles
    into four: +2 faces, +1 vertex, the mesh stays a valid manifold).  Used for BASELINE config C's "10k-face
    mesh": Spot (5856 F) + 2072 split edges = 10 000 faces."""
    v = [p for p in np.asarray(verts, dtype=np.float64)]
    f = [tuple(int(i) for i in t) for t in np.asarray(faces)]
    need = target_faces - len(f)
    if need < 0 or need % 2:
        raise ValueError("target_faces must exceed the mesh's face count by an even number")
    while need > 0:
        edge_tris = {}
        for ti, (a, b, c) in enumerate(f):
            for e in ((a, b), (b, c), (c, a)):
                edge_tris.setdefault((min(e), max(e)), []).append(ti)
        cand = [(float(np.sum((v[a] - v[b]) ** 2)), (a, b)) for (a, b), ts in edge_tris.items() if len(ts) == 2]
        cand.sort(key=lambda x: (-x[0], x[1]))
        used, split = set(), 0
        out = list(f)
        for _, (a, b) in cand:
            if need == 0:
                break
            t0, t1 = edge_tris[(a, b)]
            if t0 in used or t1 in used:
                continue
            used.update((t0, t1))
            m = len(v)
            v.append(0.5 * (v[a] + v[b]))
            for ti in (t0, t1):
                tri = f[ti]
                k = [i for i in range(3) if tri[i] not in (a, b)][0]      # the corner opposite the edge
                p, q, r = tri[k], tri[(k + 1) % 3], tri[(k + 2) % 3]      # winding kept: (p,q,r) -> (p,q,m) + (p,m,r)
                out[ti] = (p, q, m)
                out.append((p, m, r))
            need -= 2
            split += 1
        f = out
        if split == 0:
            raise ValueError("no splittable edge left")
    return np.asarray(v, dtype=np.float32), np.asarray(f, dtype=np.int32)


def mesh_by_name(name: str):
    """'spot' (2930 V / 5856 F), 'spot10k' (Spot refined to 10 000 faces, BASELINE config C), 'icoK' (icosphere level K,
    radius 0.5: ico3 642 / 1280 -- what LASR trains at --, ico4 2562 / 5120, ico5 10242 / 20480)."""
    if name == "spot":
        return load_spot()
    if name == "spot10k":
        v, f = load_spot()
        return refine_to_faces(v, f, 10000)
    if name.startswith("ico"):
        return icosphere(int(name[3:]), 0.5)
    raise ValueError(f"unknown mesh {name!r}")


def quaternion_to_matrix_np(q):
    """XYZW unit-quaternion matrix (the convention of nnutils/mesh_net.py:118,120), numpy."""
    q = np.asarray(q, dtype=np.float64)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    m = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                  2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                  2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)
    return m.reshape(q.shape[:-1] + (3, 3)).astype(np.float32)


def region_workload(mesh: str, n_frames: int, n_bones: int, image_size: int, seed: int = 0):
    """Host-side (numpy) inputs of one batch of the whole render region of LASR.forward (mesh_net.py:255-447) for
    bench.py: ``n_frames`` = 2*bs frames laid out [first frames; second frames] on the render_syn.py camera orbit,
    ``n_bones`` bones (1 body + parts with small random articulations, control points at sampled vertices) and random
    observations in the formats LASR's loader decodes them from (uint8 frames and silhouettes, float flow / occlusion).
    Shared parameters (mean shape, vertex colours, bone parameters) come separately from per-batch inputs."""
    if n_frames % 2:
        raise ValueError("the region works on frame pairs: n_frames must be even")
    rng = np.random.RandomState(100 + seed)
    v, f = mesh_by_name(mesh)
    V, N, B, S = len(v), n_frames, n_bones, image_size
    Rb, Tb, _, _ = spot_cameras(N, seed)
    q = rng.randn(N * B, 4) * 0.05
    q[:, 3] = 1
    R = quaternion_to_matrix_np(q).reshape(N, B, 3, 3)
    R[:, 0] = Rb
    T = (rng.randn(N, B, 3) * 0.01).astype(np.float32)
    T[:, 0] = Tb[:, 0]
    params = dict(mean_v=v, tex=vertex_colors(V, 0))
    if B > 1:
        idx = np.random.RandomState(7).permutation(V)[:B - 1]      # the shared parameters do not depend on the rank
        params.update(ctl_ts=v[idx].copy(), rest_ts=v[idx].copy(),
                      ctl_rs=np.tile(np.array([[0, 0, 0, 1]], np.float32), (B - 1, 1)),
                      log_ctl=np.ones((B - 1, 3), np.float32))
    batch = dict(imgs_u8=rng.randint(0, 256, (N, 3, S, S)).astype(np.uint8),
                 masks_u8=(rng.rand(N, S, S) > 0.5).astype(np.uint8),
                 flow=(rng.randn(N, 3, S, S) * 0.02).astype(np.float32),
                 occ=rng.randn(N, S, S).astype(np.float32))
    pose = dict(Rmat=np.ascontiguousarray(R.reshape(-1, 3, 3)), Tmat=np.ascontiguousarray(T.reshape(-1, 3)),
                ppoint=np.zeros((N, 2), np.float32), scale=np.full((N, 1), 10.0, np.float32))
    return dict(faces=f, params=params, pose=pose, batch=batch, V=V, F=len(f), N=N, B=B, S=S)
