"""Dense geometry helpers of SoftRas with the reference's signatures
(third_party/softras/soft_renderer/functional/{face_vertices,vertex_normals,look_at,look,
orthogonal,perspective,get_points_from_angles,ambient_lighting,directional_lighting}.py).

Plain PyTorch: these are small, bandwidth-trivial ops; the LASR fast path bypasses most of them
(see lasr_b200/soft_renderer/renderer.py)."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def face_vertices(vertices, faces):
    """[bs, nv, C] gathered by [bs, nf, 3] -> [bs, nf, 3, C] (face_vertices.py:4-22)."""
    assert vertices.ndimension() == 3
    assert faces.ndimension() == 3
    assert vertices.shape[0] == faces.shape[0]
    assert vertices.shape[2] == 3
    assert faces.shape[2] == 3
    bs, nv = vertices.shape[:2]
    offs = (torch.arange(bs, device=vertices.device, dtype=torch.int64) * nv)[:, None, None]
    return vertices.reshape(bs * nv, vertices.shape[2])[faces.long() + offs]


def vertex_normals(vertices, faces):
    """Area-weighted vertex normals, F.normalize eps=1e-6 (vertex_normals.py:4-37)."""
    assert vertices.ndimension() == 3
    assert faces.ndimension() == 3
    assert vertices.shape[0] == faces.shape[0]
    assert vertices.shape[2] == 3
    assert faces.shape[2] == 3
    bs, nv = vertices.shape[:2]
    offs = (torch.arange(bs, device=vertices.device, dtype=torch.int64) * nv)[:, None, None]
    idx = (faces.long() + offs).view(-1, 3)
    flat = vertices.reshape(bs * nv, 3)
    vf = flat[idx]
    normals = torch.zeros(bs * nv, 3, device=vertices.device, dtype=vertices.dtype)
    normals.index_add_(0, idx[:, 1], torch.cross(vf[:, 2] - vf[:, 1], vf[:, 0] - vf[:, 1], dim=1))
    normals.index_add_(0, idx[:, 2], torch.cross(vf[:, 0] - vf[:, 2], vf[:, 1] - vf[:, 2], dim=1))
    normals.index_add_(0, idx[:, 0], torch.cross(vf[:, 1] - vf[:, 0], vf[:, 2] - vf[:, 0], dim=1))
    return F.normalize(normals, eps=1e-6, dim=1).reshape(bs, nv, 3)


_const_cache = {}


def _const(values, device):
    """Small constant vectors (eye / at / up / colours given as Python lists) are uploaded once per
    (value, device) and reused: no host->device copy inside the step, so it can be CUDA-graph captured."""
    key = (tuple(float(v) for v in values), str(device))
    t = _const_cache.get(key)
    if t is None:
        t = torch.tensor(list(key[0]), dtype=torch.float32, device=device)
        _const_cache[key] = t
    return t


def _as_batch(x, batch_size, device):
    if isinstance(x, (list, tuple)):
        x = _const(x, device)
    elif isinstance(x, np.ndarray):
        x = torch.from_numpy(x).to(device)
    elif torch.is_tensor(x):
        x = x.to(device)
    if x.ndimension() == 1:
        x = x[None, :].repeat(batch_size, 1)
    return x


def look_at(vertices, eye, at=[0, 0, 0], up=[0, 1, 0]):
    """look_at.py:6-62 (axes normalised with eps=1e-5)."""
    if vertices.ndimension() != 3:
        raise ValueError("vertices Tensor should have 3 dimensions")
    bs, device = vertices.shape[0], vertices.device
    key = None
    if all(isinstance(t, (list, tuple)) for t in (eye, at, up)):
        # a camera given as Python constants (LASR: renderer.transform.transformer._eye) has a constant
        # basis: build it once per (camera, batch size, device) with the very same ops, then reuse
        key = (tuple(map(float, eye)), tuple(map(float, at)), tuple(map(float, up)), bs, str(device))
        hit = _look_at_cache.get(key)
        if hit is not None:
            eye_b, rt, identity = hit
            # LASR's camera sits on the -z axis looking at the origin (mesh_net.py:132-149): the basis is exactly the
            # identity, and x*1 + y*0 + z*0 is x bit for bit -- skip the batched 3x3 matmul and its backward
            return vertices - eye_b if identity else torch.matmul(vertices - eye_b, rt)
    eye, at, up = _as_batch(eye, bs, device), _as_batch(at, bs, device), _as_batch(up, bs, device)
    z_axis = F.normalize(at - eye, eps=1e-5)
    x_axis = F.normalize(torch.cross(up, z_axis, dim=1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=1), eps=1e-5)
    r = torch.stack((x_axis, y_axis, z_axis), dim=1)
    if vertices.shape != eye.shape:
        eye = eye[:, None, :]
    if key is not None and not (vertices.is_cuda and torch.cuda.is_current_stream_capturing()):
        if len(_look_at_cache) > 32:
            _look_at_cache.clear()
        rt = r.transpose(1, 2).contiguous()
        identity = bool((rt == torch.eye(3, dtype=rt.dtype, device=rt.device)).all())     # one sync, once per camera
        _look_at_cache[key] = (eye.contiguous(), rt, identity)
    return torch.matmul(vertices - eye, r.transpose(1, 2))


_look_at_cache = {}


def look(vertices, eye, direction=[0, 1, 0], up=None):
    """look.py: camera at `eye` looking along `direction`."""
    if vertices.ndimension() != 3:
        raise ValueError("vertices Tensor should have 3 dimensions")
    bs, device = vertices.shape[0], vertices.device
    if up is None:
        up = torch.tensor([0, 1, 0], dtype=torch.float32, device=device)
    eye = _as_batch(eye, 1, device) if not torch.is_tensor(eye) or eye.ndimension() == 1 else eye.to(device)
    direction = _as_batch(direction, 1, device)
    up = _as_batch(up, 1, device)
    z_axis = F.normalize(direction, eps=1e-5)
    x_axis = F.normalize(torch.cross(up, z_axis, dim=1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=1), eps=1e-5)
    r = torch.stack((x_axis, y_axis, z_axis), dim=1)
    if vertices.shape != eye.shape:
        eye = eye[:, None, :]
    return torch.matmul(vertices - eye, r.transpose(1, 2))


def orthogonal(vertices, scale):
    """orthogonal.py:4-17 -- x,y scaled, z kept."""
    if vertices.ndimension() != 3:
        raise ValueError("vertices Tensor should have 3 dimensions")
    if isinstance(scale, (int, float)) and scale == 1 and vertices.shape[2] == 3:
        return vertices          # x*1 is exact: value-identical to the stack below (LASR: viewing_scale=1)
    return torch.stack((vertices[:, :, 0] * scale, vertices[:, :, 1] * scale, vertices[:, :, 2]), dim=2)


def perspective(vertices, angle=30.0):
    """perspective.py -- x,y divided by z*tan(angle)."""
    if vertices.ndimension() != 3:
        raise ValueError("vertices Tensor should have 3 dimensions")
    angle = torch.tensor(angle / 180 * math.pi, dtype=torch.float32, device=vertices.device)[None]
    width = torch.tan(angle)[:, None]
    z = vertices[:, :, 2]
    return torch.stack((vertices[:, :, 0] / z / width, vertices[:, :, 1] / z / width, z), dim=2)


def get_points_from_angles(distance, elevation, azimuth, degrees=True):
    """get_points_from_angles.py -- spherical camera position."""
    if isinstance(distance, (float, int)):
        if degrees:
            elevation, azimuth = math.radians(elevation), math.radians(azimuth)
        return (distance * math.cos(elevation) * math.sin(azimuth), distance * math.sin(elevation),
                -distance * math.cos(elevation) * math.cos(azimuth))
    if degrees:
        elevation = math.pi / 180.0 * elevation
        azimuth = math.pi / 180.0 * azimuth
    return torch.stack([distance * torch.cos(elevation) * torch.sin(azimuth), distance * torch.sin(elevation),
                        -distance * torch.cos(elevation) * torch.cos(azimuth)]).transpose(1, 0)


def _color(c, device):
    if isinstance(c, (tuple, list)):
        c = _const(c, device)
    elif isinstance(c, np.ndarray):
        c = torch.from_numpy(c).float().to(device)
    return c[None, :] if c.ndimension() == 1 else c


def ambient_lighting(light, light_intensity=0.5, light_color=(1, 1, 1)):
    """ambient_lighting.py:7-18 (in place on `light`, as the reference)."""
    light += light_intensity * _color(light_color, light.device)[:, None, :]
    return light


def directional_lighting(light, normals, light_intensity=0.5, light_color=(1, 1, 1), light_direction=(0, 1, 0)):
    """directional_lighting.py:7-27."""
    color = _color(light_color, light.device)
    direction = _color(light_direction, light.device)
    cosine = F.relu(torch.sum(normals * direction, dim=2))
    light += light_intensity * (color[:, None, :] * cosine[:, :, None])
    return light


def projection(vertices, P, dist_coeffs, orig_size):
    """Pinhole projection with OpenCV's five-coefficient lens distortion (k1, k2, p1, p2, k3), then normalisation of
    the pixel coordinates to [-1, 1] by ``orig_size`` (functional/projection.py:4-32).  ``P`` is [N,3,4]; depth is
    passed through.  The operation order -- radius by sqrt, its even powers by ``**`` -- is the reference's, so the
    two agree to the last bit on the same device."""
    hom = torch.cat([vertices, torch.ones_like(vertices[:, :, :1])], dim=-1)
    cam = torch.bmm(hom, P.transpose(2, 1))
    depth = cam[:, :, 2]
    u = cam[:, :, 0] / (depth + 1e-5)
    v = cam[:, :, 1] / (depth + 1e-5)
    k1, k2, p1, p2, k3 = (dist_coeffs[:, None, i] for i in range(5))
    r = torch.sqrt(u ** 2 + v ** 2)
    radial = 1 + k1 * (r ** 2) + k2 * (r ** 4) + k3 * (r ** 6)
    ud = u * radial + 2 * p1 * u * v + p2 * (r ** 2 + 2 * u ** 2)
    vd = v * radial + p1 * (r ** 2 + 2 * v ** 2) + 2 * p2 * u * v
    to_ndc = lambda t: 2 * (t - orig_size / 2.) / orig_size
    return torch.stack([to_ndc(ud), to_ndc(vd), depth], dim=-1)
