"""The reference's training augmentation (starcop/data/datamodule.py:128-134):

    K.AugmentationSequential(K.RandomRotation(p=0.5, degrees=90), K.RandomHorizontalFlip(p=0.5),
                             K.RandomVerticalFlip(p=0.5), keepdim=True, data_keys=["input", "mask", "input"])

runs per sample on the CPU inside the DataLoader workers -- the reference's real bottleneck.  Here the
same three operators are drawn per sample on the host (a ``torch.Generator``: reproducible) and applied on the GPU
to the whole batch by ONE affine resampling (``sc_affine_warp``) per tensor: rotation by angle ~ U(-90, 90) degrees
about the image centre ((W-1)/2, (H-1)/2) with bilinear interpolation and zero fill -- kornia.warp_affine
(align_corners=True) **[3P, from memory]** -- composed with the two exact flips.  Input, label and loss weight get the
same geometry; kornia 0.6.7 resamples masks like images (bilinear), ``mask_mode="nearest"`` keeps labels binary."""
import math

import torch

from . import _lib


def draw_params(batch_size, generator=None, p_rot=0.5, degrees=90.0, p_hflip=0.5, p_vflip=0.5):
    """per-sample draws in the reference's operator order -> dict of CPU tensors (angle in degrees, 0 = identity)"""
    g = generator
    rot = torch.rand(batch_size, generator=g) < p_rot
    angle = (torch.rand(batch_size, generator=g) * 2 - 1) * degrees
    return {"angle": torch.where(rot, angle, torch.zeros_like(angle)),
            "hflip": torch.rand(batch_size, generator=g) < p_hflip,
            "vflip": torch.rand(batch_size, generator=g) < p_vflip}


def dst_to_src_matrices(params, H, W):
    """(B, 6) float32 matrices mapping an OUTPUT pixel to its source position for out = vflip(hflip(rotate(in))).
    Rotation: kornia.get_rotation_matrix2d(center, angle, 1) = [[a, b, (1-a)cx - b cy], [-b, a, b cx + (1-a) cy]],
    a = cos, b = sin (source -> rotated); its inverse is the same form with -angle."""
    B = params["angle"].numel()
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    out = torch.empty(B, 6, dtype=torch.float64)
    for i in range(B):
        th = math.radians(float(params["angle"][i]))
        a, b = math.cos(th), -math.sin(th)                      # inverse rotation
        m = [[a, b, (1 - a) * cx - b * cy], [-b, a, b * cx + (1 - a) * cy]]
        # the flips act on the OUTPUT coordinates first (they are applied after the rotation)
        if bool(params["hflip"][i]):                             # x -> W-1-x
            m = [[-r[0], r[1], r[2] + r[0] * (W - 1)] for r in m]
        if bool(params["vflip"][i]):                             # y -> H-1-y
            m = [[r[0], -r[1], r[2] + r[1] * (H - 1)] for r in m]
        out[i] = torch.tensor(m[0] + m[1], dtype=torch.float64)
    return out.float()


def affine_warp(x, mats, nearest=False):
    """x: (B,C,H,W) CUDA float32, mats: (B,6) -> resampled tensor (new buffer)."""
    if not x.is_cuda:
        raise _lib.StarcopB200Error("starcop_b200 runs on CUDA tensors only (no CPU path)")
    x = x.contiguous().float()
    B, C, H, W = x.shape
    m = mats.to(x.device, torch.float32).contiguous()
    out = torch.empty_like(x)
    _lib.call("sc_affine_warp", x.data_ptr(), out.data_ptr(), m.data_ptr(), B, C, H, W, int(nearest),
              torch.cuda.current_stream(x.device).cuda_stream)
    return out


class TrainAugmentation:
    """callable batch -> batch: the reference's spatial augmentation for the keys a training batch holds"""

    def __init__(self, seed=0, mask_mode="bilinear", p_rot=0.5, degrees=90.0, p_hflip=0.5, p_vflip=0.5):
        self.generator = torch.Generator().manual_seed(seed)
        self.mask_mode = mask_mode
        self.kw = dict(p_rot=p_rot, degrees=degrees, p_hflip=p_hflip, p_vflip=p_vflip)

    def __call__(self, batch):
        x = batch["input"]
        B, _, H, W = x.shape
        params = draw_params(B, self.generator, **self.kw)
        mats = dst_to_src_matrices(params, H, W)
        out = dict(batch)
        out["input"] = affine_warp(x, mats)
        out["output"] = affine_warp(batch["output"], mats, nearest=self.mask_mode == "nearest")
        if "weight_loss" in batch:
            out["weight_loss"] = affine_warp(batch["weight_loss"], mats)
        out["_augmentation"] = params
        return out
