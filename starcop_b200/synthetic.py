"""Deterministic synthetic inputs of the shapes BASELINE.json names.
There is no network and no dataset: every benchmark / parity input is generated here.
Host-side numpy only (the reference's data side is numpy too, dataset.py:40-102).
"""
import numpy as np
import torch

AVIRIS_WINDOW_BANDS = 73          # 2124..2485 nm on the 5 nm grid (process_aviris.py:192-195)


def _blobs(rng, h, w, n, rmin=6, rmax=30):
    yy, xx = np.mgrid[0:h, 0:w]
    m = np.zeros((h, w), dtype=bool)
    for _ in range(n):
        cy, cx = rng.integers(0, h), rng.integers(0, w)
        ry, rx = rng.integers(rmin, rmax), rng.integers(rmin, rmax)
        m |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
    return m


def _smooth_field(rng, h, w, lo, hi, cells=8):
    g = rng.uniform(lo, hi, size=(cells + 1, cells + 1))
    ys, xs = np.linspace(0, cells, h), np.linspace(0, cells, w)
    y0, x0 = np.minimum(ys.astype(int), cells - 1), np.minimum(xs.astype(int), cells - 1)
    fy, fx = (ys - y0)[:, None], (xs - x0)[None, :]
    a, b = g[y0][:, x0], g[y0][:, x0 + 1]
    c, d = g[y0 + 1][:, x0], g[y0 + 1][:, x0 + 1]
    return (a * (1 - fy) * (1 - fx) + b * (1 - fy) * fx + c * fy * (1 - fx) + d * fy * fx)


def hyperstarcop_batch(batch_size, size=512, seed=0, channels=4):
    """The DataLoader batch dict of dataset.py:59-102 for mag1c+RGB tiles (cfg 1/2/4):
    input (B,C,H,W) raw products, output (B,1,H,W) in {0,1}, weight_loss, id, has_plume."""
    rng = np.random.default_rng(seed)
    H = W = size
    x = np.empty((batch_size, channels, H, W), np.float32)
    y = np.zeros((batch_size, 1, H, W), np.float32)
    for b in range(batch_size):
        mag = np.clip(rng.exponential(150.0, size=(H, W)), 0, 10000)
        if b % 2 == 0:                                   # every other tile holds plumes
            m = _blobs(rng, H, W, int(rng.integers(1, 4)), max(2, size // 85), max(4, size // 17))
            mag = np.where(m, mag + rng.uniform(500, 3000, size=(H, W)), mag)
            y[b, 0] = m
        x[b, 0] = np.clip(mag, 0, 10000)
        for c in range(1, channels):
            x[b, c] = _smooth_field(rng, H, W, 5.0, 60.0)
    w = np.clip(x[:, 0:1] / 400, 0.1, 1).astype(np.float32)      # feature_extration.py:32-35
    n_px = 10 * H * W / 64 ** 2
    return {"input": torch.from_numpy(x), "output": torch.from_numpy(y),
            "weight_loss": torch.from_numpy(w),
            "id": [f"synthetic_{seed}_{b}" for b in range(batch_size)],
            "has_plume": torch.from_numpy((y.sum((1, 2, 3)) > n_px).astype(np.int64))}


def synthetic_template(n_bands=AVIRIS_WINDOW_BANDS, seed=7):
    """A CH4-like unit absorption spectrum (negative, band structured) when the real
    ``generate_template_from_bands`` output (tests/golden/ch4_template_aviris.npz) is not at hand."""
    k = np.arange(n_bands)
    t = -(0.25 + 0.9 * np.exp(-((k - 0.62 * n_bands) / (0.09 * n_bands)) ** 2)
          + 0.45 * np.abs(np.sin(k * 0.9)) * np.exp(-((k - 0.5 * n_bands) / (0.35 * n_bands)) ** 2))
    return t.astype(np.float64)


def aviris_cube(n_tiles=1, size=512, bands=125, window=(52, 125), seed=0, dtype=np.float32, template=None):
    """(n, H, W, C) BIP radiance cube (process_aviris.py:183-184 opens interleave='bip'):
    albedo(h,w) * mu(c) * (1 + alpha(h,w) * t(c)) + noise; SWIR window = bands[window[0]:window[1]]."""
    rng = np.random.default_rng(seed)
    H = W = size
    nwin = window[1] - window[0]
    t_win = synthetic_template(nwin) if template is None else np.asarray(template, np.float64)
    t = np.zeros(bands); t[window[0]:window[1]] = t_win
    c = np.arange(bands)
    mu = 8.0 * np.exp(-c / (0.9 * bands)) + 0.6 + 0.15 * np.sin(c * 0.37)
    cube = np.empty((n_tiles, H, W, bands), dtype)
    alpha = np.zeros((n_tiles, H, W))
    for n in range(n_tiles):
        albedo = _smooth_field(rng, H, W, 0.5, 1.5)
        m = _blobs(rng, H, W, 2, max(2, size // 85), max(4, size // 17))
        alpha[n] = np.where(m, rng.uniform(0.002, 0.03, size=(H, W)), 0.0)
        rad = albedo[..., None] * mu * (1.0 + alpha[n][..., None] * t)
        rad = rad + rng.normal(0.0, 0.01, size=rad.shape) * mu
        cube[n] = rad.astype(dtype)
    return cube, t_win.astype(dtype), alpha


def ratio_bands(size=512, seed=0):
    """Two positive float32 bands (bg, sig) with a few exact zeros (nodata corner)."""
    rng = np.random.default_rng(seed)
    bg = (_smooth_field(rng, size, size, 0.5, 3.0) + rng.normal(0, 0.02, (size, size))).astype(np.float32)
    sig = (bg * 0.8 + rng.normal(0, 0.02, (size, size))).astype(np.float32)
    bg[: size // 16, : size // 16] = 0
    sig[: size // 16, : size // 16] = 0
    return np.abs(bg), np.abs(sig)
