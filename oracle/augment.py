"""Restatement of the reference's training augmentation (starcop/data/datamodule.py:128-134, kornia 0.6.7
``RandomRotation`` / ``RandomHorizontalFlip`` / ``RandomVerticalFlip``) with torch CPU primitives: the rotation is
``F.grid_sample(align_corners=True, padding_mode="zeros")`` on the affine grid of ``get_rotation_matrix2d`` -- what
``kornia.geometry.transform.warp_affine`` does **[3P, from memory]** -- and the flips are ``torch.flip``.
Test infrastructure only (see ``oracle/__init__.py``)."""
import math

import torch
import torch.nn.functional as F


def rotate(x, angle_deg, mode="bilinear"):
    """x: (C,H,W); rotation about ((W-1)/2, (H-1)/2), positive angle = kornia / OpenCV convention."""
    C, H, W = x.shape
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    th = math.radians(angle_deg)
    a, b = math.cos(th), math.sin(th)
    M = torch.tensor([[a, b, (1 - a) * cx - b * cy], [-b, a, b * cx + (1 - a) * cy], [0, 0, 1]], dtype=torch.float64)
    Minv = torch.linalg.inv(M)                                     # rotated (dst) -> source, pixel coordinates
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    sx = Minv[0, 0] * xs + Minv[0, 1] * ys + Minv[0, 2]
    sy = Minv[1, 0] * xs + Minv[1, 1] * ys + Minv[1, 2]
    grid = torch.stack([2 * sx / (W - 1) - 1, 2 * sy / (H - 1) - 1], dim=-1)[None].float()
    return F.grid_sample(x[None].float(), grid, mode=mode, padding_mode="zeros", align_corners=True)[0]


def augment_sample(x, angle_deg, hflip, vflip, mode="bilinear"):
    out = rotate(x, angle_deg, mode) if angle_deg != 0 else x.float()
    if hflip:
        out = torch.flip(out, dims=(-1,))
    if vflip:
        out = torch.flip(out, dims=(-2,))
    return out
