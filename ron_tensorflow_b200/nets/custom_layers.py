"""Drop-in for the one hot-path function of the reference's ``nets/custom_layers.py``:
``modified_smooth_l1`` (:31-50), element-wise on CUDA tensors (``ronk_smooth_l1``).  The layer
definitions of that file (pad2d, l2_normalization, ...) belong to the network and are out of scope."""
from .. import core

__all__ = ['modified_smooth_l1']


def modified_smooth_l1(bbox_pred, bbox_targets, bbox_inside_weights=1., bbox_outside_weights=1., sigma=1.):
    """reference nets/custom_layers.py:31-50:
    ResultLoss = outside_weights * SmoothL1(inside_weights * (bbox_pred - bbox_targets)),
    SmoothL1(x) = 0.5 * (sigma * x)^2 if |x| < 1 / sigma^2 else |x| - 0.5 / sigma^2.
    The weights are scalars (the reference only ever passes the defaults)."""
    return core.smooth_l1(bbox_pred, bbox_targets, bbox_inside_weights, bbox_outside_weights, sigma)
