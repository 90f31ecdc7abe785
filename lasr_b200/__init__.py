"""lasr_b200 -- B200-native (sm_100a) implementation of LASR's per-frame differentiable render
hot path: LBS skinning + camera projection -> SoftRas soft-rasterize fwd/bwd -> image losses.
See DESIGN.md for the scope table and INTEGRATION.md for how the reference binds to it."""
import sys as _sys

__all__ = ["install_as_soft_renderer"]


def install_as_soft_renderer():
    """Make ``import soft_renderer as sr`` resolve to lasr_b200.soft_renderer (and its
    ``functional`` submodule), so the reference's nnutils/*.py run unchanged."""
    from . import soft_renderer as _sr
    _sys.modules["soft_renderer"] = _sr
    _sys.modules["soft_renderer.functional"] = _sr.functional
    # the brute-force nearest-neighbour extension LASR imports next to it (nnutils/train_utils.py:43)
    from . import chamfer3D as _ch
    _sys.modules["chamfer3D"] = _ch
    _sys.modules["chamfer3D.dist_chamfer_3D"] = _ch.dist_chamfer_3D
    return _sr
