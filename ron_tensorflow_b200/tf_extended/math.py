"""reference: tf_extended/math.py:25-67 (safe_divide, cummax).  Metric-side helpers."""
import torch

__all__ = ['safe_divide', 'cummax']


def safe_divide(numerator, denominator, name=None):
    """reference :25-38: 0 where the denominator is <= 0."""
    num = torch.as_tensor(numerator)
    den = torch.as_tensor(denominator, dtype=num.dtype if num.is_floating_point() else None, device=num.device)
    return torch.where(den > 0, num / torch.where(den > 0, den, torch.ones_like(den)), torch.zeros_like(num))


def cummax(x, reverse=False, name=None):
    """reference :41-67: cumulative maximum of a 1-D tensor."""
    x = torch.as_tensor(x)
    if reverse:
        return torch.flip(torch.cummax(torch.flip(x, [0]), 0).values, [0])
    return torch.cummax(x, 0).values
