"""``Mesh`` container with the reference's surface (third_party/softras/soft_renderer/mesh.py:9-178).

Holds batched vertices / faces / textures and lazily derives the gathered per-face tensors the
rasterizer consumes.  OBJ I/O and voxelisation are outside the hot path."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import functional as srf


class Mesh(object):
    def __init__(self, vertices, faces, textures=None, texture_res=1, texture_type="surface", shared_faces=None):
        # shared_faces (extension): an int32 [F,3] topology the caller guarantees every frame uses; lets the
        # rasterizer take the indexed path without comparing the frames' face tensors (a device sync).
        self.shared_faces = shared_faces
        if isinstance(vertices, np.ndarray):
            vertices = torch.from_numpy(vertices).float().cuda()
        if isinstance(faces, np.ndarray):
            faces = torch.from_numpy(faces).int().cuda()
        if vertices.ndimension() == 2:
            vertices = vertices[None]
        if faces.ndimension() == 2:
            faces = faces[None]
        self._vertices = vertices
        self._faces = faces
        self.device = vertices.device
        self.texture_type = texture_type
        self.batch_size = vertices.shape[0]
        self.num_vertices = vertices.shape[1]
        self.num_faces = faces.shape[1]
        self._face_vertices = None
        self._surface_normals = None
        self._vertex_normals = None
        self._fill_back = False

        if textures is None:
            if texture_type == "surface":
                textures = torch.ones(self.batch_size, self.num_faces, texture_res ** 2, 3, dtype=torch.float32,
                                      device=self.device)
                self.texture_res = texture_res
            elif texture_type == "vertex":
                textures = torch.ones(self.batch_size, self.num_vertices, 3, dtype=torch.float32, device=self.device)
                self.texture_res = 1
        else:
            if isinstance(textures, np.ndarray):
                textures = torch.from_numpy(textures).float().cuda()
            if textures.ndimension() == 3 and texture_type == "surface":
                textures = textures[None]
            if textures.ndimension() == 2 and texture_type == "vertex":
                textures = textures[None]
            self.texture_res = int(np.sqrt(textures.shape[2]))
        self._textures = textures
        self._origin_vertices, self._origin_faces, self._origin_textures = vertices, faces, textures

    def _invalidate(self):
        self._face_vertices = None
        self._surface_normals = None
        self._vertex_normals = None

    @property
    def faces(self):
        return self._faces

    @faces.setter
    def faces(self, faces):
        self.shared_faces = None
        self._faces = faces
        self.num_faces = faces.shape[1]
        self._invalidate()

    @property
    def vertices(self):
        return self._vertices

    @vertices.setter
    def vertices(self, vertices):
        self._vertices = vertices
        self.num_vertices = vertices.shape[1]
        self._invalidate()

    @property
    def textures(self):
        return self._textures

    @textures.setter
    def textures(self, textures):
        self._textures = textures

    @property
    def face_vertices(self):
        if self._face_vertices is None:
            self._face_vertices = srf.face_vertices(self.vertices, self.faces)
        return self._face_vertices

    @property
    def surface_normals(self):
        if self._surface_normals is None:
            fv = self.face_vertices
            v10 = fv[:, :, 0] - fv[:, :, 1]
            v12 = fv[:, :, 2] - fv[:, :, 1]
            self._surface_normals = F.normalize(torch.cross(v12, v10, dim=2), p=2, dim=2, eps=1e-6)
        return self._surface_normals

    @property
    def vertex_normals(self):
        if self._vertex_normals is None:
            self._vertex_normals = srf.vertex_normals(self.vertices, self.faces)
        return self._vertex_normals

    @property
    def face_textures(self):
        if self.texture_type in ["surface"]:
            return self.textures
        elif self.texture_type in ["vertex"]:
            return srf.face_vertices(self.textures, self.faces)
        raise ValueError("texture type not applicable")

    def fill_back_(self):
        if not self._fill_back:
            self.faces = torch.cat((self.faces, self.faces[:, :, [2, 1, 0]]), dim=1)
            self.textures = torch.cat((self.textures, self.textures), dim=1)
            self._fill_back = True

    def reset_(self):
        self.vertices = self._origin_vertices
        self.faces = self._origin_faces
        self.textures = self._origin_textures
        self._fill_back = False

    @classmethod
    def from_obj(cls, filename_obj, normalization=False, load_texture=False, texture_res=1, texture_type="surface"):
        """A Mesh from a Wavefront .obj file; with ``load_texture`` also its 'surface' texture atlas (sampled by
        ``lasr_b200_load_textures``) or its per-vertex colours (reference: mesh.py:148-164)."""
        loaded = srf.load_obj(filename_obj, normalization=normalization, texture_res=texture_res,
                              load_texture=bool(load_texture),
                              **({"texture_type": texture_type} if load_texture else {}))
        vertices, faces = loaded[0], loaded[1]
        textures = loaded[2] if load_texture else None
        return cls(vertices, faces, textures, texture_res, texture_type)

    def save_obj(self, filename_obj, save_texture=False, texture_res_out=16):
        """Write the (single) mesh of this container; ``save_texture`` adds the .mtl and the .png atlas painted by
        ``lasr_b200_create_texture_image`` (reference: mesh.py:166-174)."""
        if self.batch_size != 1:
            raise ValueError("Could not save when batch size >= 1")
        extra = dict(textures=self.textures[0], texture_res=texture_res_out, texture_type=self.texture_type) \
            if save_texture else dict(textures=None)
        srf.save_obj(filename_obj, self.vertices[0], self.faces[0], **extra)

    def voxelize(self, voxel_size=32):
        srf.voxelization(None)
