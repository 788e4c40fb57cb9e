"""Tile-index arithmetic and whole-scene prediction helpers (host side, integer exact):
``tiled_dataframe`` windows / ids / has_plume (starcop/data/datamodule.py:17-64 on georeader's
``create_windows``) and ``padded_predict`` (starcop/models/utils/padding.py:5-50)."""
import numpy as np
import torch


def create_windows(shape, window_size, overlap, include_incomplete=False):
    """georeader.slices.create_windows as the reference calls it (datamodule.py:27-28):
    row-major windows (row_off, col_off, height, width), step = window_size - overlap."""
    step_r, step_c = window_size[0] - overlap[0], window_size[1] - overlap[1]
    out = []
    for r in range(0, shape[0], step_r):
        for c in range(0, shape[1], step_c):
            h, w = min(window_size[0], shape[0] - r), min(window_size[1], shape[1] - c)
            if not include_incomplete and (h < window_size[0] or w < window_size[1]):
                continue
            out.append((r, c, h, w))
    return out


def tile_id(base_id, window):
    """datamodule.py:57-60: f"{id}_r{row_off}_c{col_off}_w{width}_h{height}"."""
    r, c, h, w = window
    return f"{base_id}_r{r}_c{c}_w{w}_h{h}"


def frac_positives(label):
    return float(torch.as_tensor(label).sum().item()) / float(np.prod(tuple(label.shape)))


def has_plume(label):
    """datamodule.py:44-50."""
    return frac_positives(label) > (10 / 64 ** 2)


def find_padding(v, divisor=8):
    """padding.py:5-10."""
    v_divisible = max(divisor, int(divisor * np.ceil(v / divisor)))
    total_pad = v_divisible - v
    pad_1 = total_pad // 2
    return pad_1, total_pad - pad_1


@torch.no_grad()
def padded_predict(tensor, model, divisor=32, device=None):
    """padding.py:13-50: reflect-pad a (C, H, W) scene to a multiple of `divisor`, run the model on
    the whole scene in one pass, crop back.  `tensor` may be a numpy array or a tensor; the padded
    scene is run on the model's device and the result returned as a numpy array like the reference."""
    assert len(tensor.shape) == 3, f"Expected 3D tensor, found {len(tensor.shape)}D tensor"
    dev = device if device is not None else next(model.parameters()).device
    t = torch.as_tensor(np.asarray(tensor) if not torch.is_tensor(tensor) else tensor).to(dev).float()
    pad_r, pad_c = find_padding(t.shape[-2], divisor), find_padding(t.shape[-1], divisor)
    padded = torch.nn.functional.pad(t[None], (pad_c[0], pad_c[1], pad_r[0], pad_r[1]), mode="reflect")
    pred = model(padded)[0]
    sr = slice(pad_r[0], None if pad_r[1] <= 0 else -pad_r[1])
    sc = slice(pad_c[0], None if pad_c[1] <= 0 else -pad_c[1])
    if pred.dim() == 3:
        pred = pred[:, sr, sc]
    elif pred.dim() == 2:
        pred = pred[sr, sc]
    else:
        raise NotImplementedError(f"Don't know how to slice the tensor of shape {pred.shape}")
    return pred.cpu().numpy().copy()
