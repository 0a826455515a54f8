"""reference: tf_extended/tensors.py:34-86 (get_shape, pad_axis).  Shape plumbing only."""
import torch
import torch.nn.functional as F

__all__ = ['get_shape', 'pad_axis']


def get_shape(x, rank=None):
    """reference :34-56: dimensions as a list of ints (always static here)."""
    shape = list(x.shape)
    if rank is not None and len(shape) != rank:
        raise ValueError('Shape %s must have rank %d' % (tuple(shape), rank))
    return shape


def pad_axis(x, offset, size, axis=0, name=None):
    """reference :59-86: zero-pad ``axis`` with ``offset`` leading zeros up to ``size``; a tensor
    already longer than ``size`` keeps its length."""
    rank = x.dim()
    axis = axis % rank
    new_size = max(int(size) - int(offset) - int(x.shape[axis]), 0)
    pad = [0, 0] * rank
    pad[2 * (rank - 1 - axis)] = int(offset)
    pad[2 * (rank - 1 - axis) + 1] = new_size
    return F.pad(x, pad, mode='constant', value=0)
