from .geometry import (ambient_lighting, directional_lighting, face_vertices, get_points_from_angles, look, look_at,
                       orthogonal, perspective, projection, vertex_normals)
from .obj_io import create_texture_image, load_mtl, load_obj, load_textures, save_obj
from .render_indexed import (PackedGradLink, RenderIndexedFunction, RenderSetsFunction, render_indexed, render_sets,
                             shared_topology, vertex_adjacency)
from .soft_rasterize import SoftRasterizeFunction, soft_rasterize


def _out_of_scope(name):
    def f(*a, **k):
        raise NotImplementedError(
            f"soft_renderer.functional.{name} is outside the LASR render hot path this package replaces "
            "(SURVEY.md section 2 rows 13/17); use the reference SoftRas for voxel tooling")
    f.__name__ = name
    return f


save_voxel = _out_of_scope("save_voxel")
voxelization = _out_of_scope("voxelization")
