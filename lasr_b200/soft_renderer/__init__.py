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

__version__ = "1.0.0+lasr_b200"
