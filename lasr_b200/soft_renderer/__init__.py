"""Drop-in for ``import soft_renderer as sr`` on the LASR render hot path.

Same public names as third_party/softras/soft_renderer/__init__.py; the rasterizer is the
sm_100a implementation in lasr_b200/csrc.  ``lasr_b200.install_as_soft_renderer()`` registers
this package under the name ``soft_renderer`` so nnutils/mesh_net.py:47, nnutils/geom_utils.py:24
and nnutils/train_utils.py:34 import it unchanged."""
from . import functional
from .lighting import AmbientLighting, DirectionalLighting, Lighting
from .mesh import Mesh
from .rasterizer import SoftRasterizer
from .renderer import SoftRenderer
from .transform import Look, LookAt, Projection, Transform

Rasterizer = SoftRasterizer  # the reference's hard `Renderer`/`Rasterizer` pair is not on the hot path
from ..loss_utils import FlattenLoss, LaplacianLoss        # soft_renderer/losses.py:6-113 (same API as ext_nnutils')
from .. import loss_utils as losses                        # `soft_renderer.losses`


class Renderer:
    """The reference's hard-rasterizing `Renderer` (soft_renderer/renderer.py:11-46) builds `sr.Rasterizer`, a name
    the reference package itself never defines -- constructing it fails there with AttributeError.  LASR never does;
    it is refused here with a message instead."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("soft_renderer.Renderer (hard rasterization) is outside the LASR render hot path; "
                                  "use SoftRenderer")

__version__ = "1.0.0+lasr_b200"
