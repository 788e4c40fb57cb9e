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


def tiled_records(records, labels, tile_size, overlap, scene_shape=(512, 512)):
    """``tiled_dataframe`` (datamodule.py:17-64) without pandas / georeader: every record (a dict with at least
    ``"id"``) is expanded into one record per window of ``create_windows(scene_shape, tile_size, overlap)``;
    ``labels[i]`` is record i's (H, W) binary label, from which each tile's ``frac_positives`` and
    ``has_plume = frac_positives > 10 / 64**2`` are computed.  The tile keeps the original id in ``id_original``
    and gets ``id = f"{id}_r{row}_c{col}_w{w}_h{h}"`` plus the four ``window_*`` fields."""
    out = []
    wins = create_windows(scene_shape, tile_size, overlap, include_incomplete=False)
    for rec, lab in zip(records, labels):
        lab = np.asarray(lab.cpu() if torch.is_tensor(lab) else lab)
        for (r, c, h, w) in wins:
            t = {k: v for k, v in rec.items() if k not in ("window_row_off", "window_col_off", "window_width", "window_height")}
            tile = lab[..., r:r + h, c:c + w]
            t["window"] = (r, c, h, w)
            t["frac_positives"] = float(tile.sum()) / float(tile.size)
            t["has_plume"] = t["frac_positives"] > (10 / 64 ** 2)
            t["window_col_off"], t["window_row_off"], t["window_width"], t["window_height"] = c, r, w, h
            t["id_original"] = rec["id"]
            t["id"] = tile_id(rec["id"], (r, c, h, w))
            out.append(t)
    return out


def add_sample_weight(has_plume_flags):
    """datamodule.py:309-315: inverse-frequency weights for the WeightedRandomSampler --
    1 / plume_fraction for plume tiles, 1 / (1 - plume_fraction) for the others (float64, like pandas)."""
    flags = np.asarray(has_plume_flags).astype(bool)
    plume_fraction = np.sum(flags) / flags.shape[0]
    with np.errstate(divide="ignore"):
        plume_weight = 1 / plume_fraction
        non_plume_weight = 1 / (1 - plume_fraction)
    return np.where(flags, plume_weight, non_plume_weight)


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
