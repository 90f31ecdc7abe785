"""Ambient + directional vertex/surface lighting with the reference's module surface
(third_party/softras/soft_renderer/lighting.py:9-67)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as srf


class AmbientLighting(nn.Module):
    def __init__(self, light_intensity=0.5, light_color=(1, 1, 1)):
        super().__init__()
        self.light_intensity = light_intensity
        self.light_color = light_color

    def forward(self, light):
        return srf.ambient_lighting(light, self.light_intensity, self.light_color)


class DirectionalLighting(nn.Module):
    def __init__(self, light_intensity=0.5, light_color=(1, 1, 1), light_direction=(0, 1, 0)):
        super().__init__()
        self.light_intensity = light_intensity
        self.light_color = light_color
        self.light_direction = light_direction

    def forward(self, light, normals):
        return srf.directional_lighting(light, normals, self.light_intensity, self.light_color, self.light_direction)


def _is_zero(x):
    return isinstance(x, (int, float)) and x == 0


class Lighting(nn.Module):
    def __init__(self, light_mode="surface", intensity_ambient=0.5, color_ambient=[1, 1, 1],
                 intensity_directionals=0.5, color_directionals=[1, 1, 1], directions=[0, 1, 0]):
        super().__init__()
        if light_mode not in ["surface", "vertex"]:
            raise ValueError("Lighting mode only support surface and vertex")
        self.light_mode = light_mode
        self.ambient = AmbientLighting(intensity_ambient, color_ambient)
        self.directionals = nn.ModuleList([DirectionalLighting(intensity_directionals, color_directionals,
                                                               directions)])

    def is_identity(self):
        """True when lighting multiplies textures by exactly 1 (LASR: ambient 1, directional 0,
        nnutils/mesh_net.py:132-149).  The reference still computes vertex normals and multiplies
        them by 0 in that case; skipping it is value-identical."""
        amb = self.ambient
        if not (isinstance(amb.light_intensity, (int, float)) and amb.light_intensity == 1):
            return False
        if not all(float(c) == 1.0 for c in amb.light_color):
            return False
        return all(_is_zero(d.light_intensity) for d in self.directionals)

    def forward(self, mesh):
        if self.is_identity():
            return mesh
        if self.light_mode == "surface":
            light = torch.zeros_like(mesh.faces, dtype=torch.float32).contiguous()
            light = self.ambient(light)
            for directional in self.directionals:
                if not _is_zero(directional.light_intensity):
                    light = directional(light, mesh.surface_normals)
            mesh.textures = mesh.textures * light[:, :, None, :]
        elif self.light_mode == "vertex":
            light = torch.zeros_like(mesh.vertices, dtype=torch.float32).contiguous()
            light = self.ambient(light)
            for directional in self.directionals:
                if not _is_zero(directional.light_intensity):
                    light = directional(light, mesh.vertex_normals)
            mesh.textures = mesh.textures * light
        return mesh
