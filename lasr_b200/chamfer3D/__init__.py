"""Mirror of the reference's ``chamfer3D`` package (third_party/chamfer3D): ``import chamfer3D.dist_chamfer_3D``
(nnutils/train_utils.py:43,123) resolves here after ``lasr_b200.install_as_soft_renderer()``."""
from . import dist_chamfer_3D  # noqa: F401
