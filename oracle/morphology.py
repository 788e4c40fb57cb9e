"""Restatement of ``starcop/baselines.py:25-27`` ``binary_opening`` (kornia 0.6.7
``erosion`` then ``dilation`` with the 3x3 cross, default ``border_type="geodesic"``:
out-of-image neighbours never erode and never dilate).

Test infrastructure only (see ``oracle/__init__.py``).
"""
import torch
import torch.nn.functional as F

CROSS = torch.tensor([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=torch.bool)


def _neigh(x, pad_value):
    xp = F.pad(x, (1, 1, 1, 1), value=pad_value)
    H, W = x.shape[-2:]
    return [xp[..., 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
            for dy, dx in ((0, 0), (-1, 0), (1, 0), (0, -1), (0, 1))]


def erosion_cross(x):
    """min over the cross; geodesic border = +inf padding (here 1 for {0,1} input)."""
    return torch.stack(_neigh(x.float(), 1.0)).amin(0)


def dilation_cross(x):
    """max over the cross; geodesic border = -inf padding (here 0 for {0,1} input)."""
    return torch.stack(_neigh(x.float(), 0.0)).amax(0)


def binary_opening(x):
    """baselines.py:25-27 with kernel = cross (baselines.py:41-43)."""
    eroded = torch.clamp(erosion_cross(x), 0, 1) > 0
    return torch.clamp(dilation_cross(eroded), 0, 1) > 0


def apply_threshold(pred, threshold):
    """baselines.py:54-58 (Mag1cBaseline; Sanchez/Varon identical with their thresholds)."""
    return binary_opening(pred > threshold).long()
