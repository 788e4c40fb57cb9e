"""Restatement of the tile-index arithmetic: ``starcop/data/datamodule.py:17-64``
(``tiled_dataframe``) on top of georeader ``slices.create_windows`` [third party, not in
the reference tree; pinned by the reference notebook's "441 chips from 9 tiles" =
49 windows per 512x512 tile for size 128 / overlap 64].

Test infrastructure only (see ``oracle/__init__.py``).
"""
import numpy as np


def create_windows(shape, window_size, overlap, include_incomplete=False):
    """-> list of (row_off, col_off, height, width), row-major, georeader order."""
    out = []
    step_r, step_c = window_size[0] - overlap[0], window_size[1] - overlap[1]
    for r in range(0, shape[0], step_r):
        for c in range(0, shape[1], step_c):
            h, w = min(window_size[0], shape[0] - r), min(window_size[1], shape[1] - c)
            if not include_incomplete and (h < window_size[0] or w < window_size[1]):
                continue
            out.append((r, c, h, w))
    return out


def tile_id(base_id, win):
    """datamodule.py:57-60."""
    r, c, h, w = win
    return f"{base_id}_r{r}_c{c}_w{w}_h{h}"


def has_plume(label):
    """datamodule.py:44-50: frac_positives > 10/64**2."""
    return (float(np.sum(label)) / np.prod(label.shape)) > (10 / 64 ** 2)


def find_padding(v, divisor=32):
    """starcop/models/utils/padding.py:5-10."""
    v_div = max(divisor, int(divisor * np.ceil(v / divisor)))
    total = v_div - v
    return total // 2, total - total // 2
