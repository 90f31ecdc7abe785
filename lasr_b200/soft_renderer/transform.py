"""Camera transforms with the reference's module surface
(third_party/softras/soft_renderer/transform.py:8-107).  LASR only uses
``camera_mode='look_at', perspective=False`` and pokes ``transform.transformer._eye``
(nnutils/mesh_net.py:81,354)."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import functional as srf


def _default_eye(viewing_angle):
    return [0, 0, -(1.0 / math.tan(math.radians(viewing_angle)) + 1)]


class Projection(nn.Module):
    """Calibrated camera: a [N,3,4] projection matrix plus OpenCV distortion coefficients (transform.py:8-26).
    numpy matrices are moved to the current CUDA device as the reference does; tensors stay where they are."""

    def __init__(self, P, dist_coeffs=None, orig_size=512):
        super().__init__()
        if isinstance(P, np.ndarray):
            P = torch.from_numpy(P).cuda()
        if P is None or P.ndimension() != 3 or P.shape[1] != 3 or P.shape[2] != 4:
            raise ValueError("You need to provide a valid (batch_size)x3x4 projection matrix")
        if dist_coeffs is None:
            dist_coeffs = torch.zeros(P.shape[0], 5, dtype=torch.float32, device=P.device)
        self.P, self.dist_coeffs, self.orig_size = P, dist_coeffs, orig_size

    def forward(self, vertices):
        return srf.projection(vertices, self.P, self.dist_coeffs, self.orig_size)


class LookAt(nn.Module):
    def __init__(self, perspective=True, viewing_angle=30, viewing_scale=1.0, eye=None):
        super().__init__()
        self.perspective = perspective
        self.viewing_angle = viewing_angle
        self.viewing_scale = viewing_scale
        self._eye = eye if eye is not None else _default_eye(viewing_angle)

    def forward(self, vertices):
        vertices = srf.look_at(vertices, self._eye)
        if self.perspective:
            return srf.perspective(vertices, angle=self.viewing_angle)
        return srf.orthogonal(vertices, scale=self.viewing_scale)


class Look(nn.Module):
    def __init__(self, camera_direction=[0, 0, 1], perspective=True, viewing_angle=30, viewing_scale=1.0, eye=None):
        super().__init__()
        self.perspective = perspective
        self.viewing_angle = viewing_angle
        self.viewing_scale = viewing_scale
        self.camera_direction = camera_direction
        self._eye = eye if eye is not None else _default_eye(viewing_angle)

    def forward(self, vertices):
        vertices = srf.look(vertices, self._eye, self.camera_direction)
        if self.perspective:
            return srf.perspective(vertices, angle=self.viewing_angle)
        return srf.orthogonal(vertices, scale=self.viewing_scale)


class Transform(nn.Module):
    def __init__(self, camera_mode="projection", P=None, dist_coeffs=None, orig_size=512, perspective=True,
                 viewing_angle=30, viewing_scale=1.0, eye=None, camera_direction=[0, 0, 1]):
        super().__init__()
        self.camera_mode = camera_mode
        if camera_mode == "projection":
            self.transformer = Projection(P, dist_coeffs, orig_size)
        elif camera_mode == "look":
            self.transformer = Look(camera_direction, perspective, viewing_angle, viewing_scale, eye)
        elif camera_mode == "look_at":
            self.transformer = LookAt(perspective, viewing_angle, viewing_scale, eye)
        else:
            raise ValueError("Camera mode has to be one of projection, look or look_at")

    def forward(self, mesh):
        mesh.vertices = self.transformer(mesh.vertices)
        return mesh

    def set_eyes_from_angles(self, distances, elevations, azimuths):
        if self.camera_mode not in ["look", "look_at"]:
            raise ValueError("Projection does not need to set eyes")
        self.transformer._eye = srf.get_points_from_angles(distances, elevations, azimuths)

    def set_eyes(self, eyes):
        if self.camera_mode not in ["look", "look_at"]:
            raise ValueError("Projection does not need to set eyes")
        self.transformer._eye = eyes

    @property
    def eyes(self):
        return self.transformer._eye
