"""``SoftRenderer`` (third_party/softras/soft_renderer/renderer.py:47-101): lighting -> camera ->
soft rasterizer, with the attribute surface LASR pokes at run time
(``renderer.rasterizer.{near,far,sigma_val,background_color}``,
``renderer.transform.transformer._eye``; nnutils/mesh_net.py:81,306-316,354-356)."""
from __future__ import annotations

import torch.nn as nn

from .lighting import Lighting
from .mesh import Mesh
from .rasterizer import SoftRasterizer
from .transform import Transform


class SoftRenderer(nn.Module):
    def __init__(self, image_size=256, background_color=[0, 0, 0], near=1, far=100, anti_aliasing=False,
                 fill_back=True, eps=1e-3, sigma_val=1e-5, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-4,
                 aggr_func_rgb="softmax", aggr_func_alpha="prod", texture_type="surface", camera_mode="projection",
                 P=None, dist_coeffs=None, orig_size=512, perspective=True, viewing_angle=30, viewing_scale=1.0,
                 eye=None, camera_direction=[0, 0, 1], light_mode="surface", light_intensity_ambient=0.5,
                 light_color_ambient=[1, 1, 1], light_intensity_directionals=0.5, light_color_directionals=[1, 1, 1],
                 light_directions=[0, 1, 0]):
        super().__init__()
        self.lighting = Lighting(light_mode, light_intensity_ambient, light_color_ambient,
                                 light_intensity_directionals, light_color_directionals, light_directions)
        self.transform = Transform(camera_mode, P, dist_coeffs, orig_size, perspective, viewing_angle, viewing_scale,
                                   eye, camera_direction)
        self.rasterizer = SoftRasterizer(image_size, background_color, near, far, anti_aliasing, fill_back, eps,
                                         sigma_val, dist_func, dist_eps, gamma_val, aggr_func_rgb, aggr_func_alpha,
                                         texture_type)

    def set_sigma(self, sigma):
        self.rasterizer.sigma_val = sigma

    def set_gamma(self, gamma):
        self.rasterizer.gamma_val = gamma

    def set_texture_mode(self, mode):
        assert mode in ["vertex", "surface"], "Mode only support surface and vertex"
        self.lighting.light_mode = mode
        self.rasterizer.texture_type = mode

    def render_mesh(self, mesh, mode=None):
        self.set_texture_mode(mesh.texture_type)
        mesh = self.lighting(mesh)
        mesh = self.transform(mesh)
        return self.rasterizer(mesh, mode)

    def forward(self, vertices, faces, textures=None, mode=None, texture_type="surface"):
        mesh = Mesh(vertices, faces, textures=textures, texture_type=texture_type)
        return self.render_mesh(mesh, mode)
