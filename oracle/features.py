"""Restatement of ``starcop/data/feature_extration.py``: ``weight_mag1c`` (:32-35),
``no_outliers`` (:37-40), ``ratio_2c_match_c_from_sums_outlier`` (:42-56),
``ratio_MLR_local`` c_matched_outliers branch (:58-112) and the EMIT rescale of
``starcop/emit_tools/emit_dataset.py:62-69,80-101``.

numpy, same calls as the reference (np.percentile linear interpolation, np.sum pairwise).
Test infrastructure only (see ``oracle/__init__.py``).
"""
import numpy as np


def weight_mag1c(mag1c):
    return np.clip(mag1c / 400, 0.1, 1)


def no_outliers(d, percentile=5):
    hi = np.percentile(d, 100 - percentile)
    lo = np.percentile(d, percentile)
    return d[np.where((d >= lo) & (d <= hi))]


def ratio_2c_match_c_from_sums_outlier(background_channel, signal, p=5, zero_value_out=-.6):
    zero = (signal < 1e-6) & (background_channel < 1e-6)
    c = np.sum(no_outliers(background_channel.flatten(), p)) / np.sum(no_outliers(signal.flatten(), p))
    R = (c * signal - background_channel) / (background_channel + 1e-6)
    R[zero] = zero_value_out
    return R


def ratio_mlr_local(bands_bg, band_target):
    """feature_extration.py:58-112, division="c_matched_outliers": sklearn LinearRegression
    (with intercept) == lstsq on the centred data; restated with numpy float64 lstsq."""
    shape = band_target.shape
    X = np.swapaxes(np.asarray([b.flatten() for b in bands_bg]), 0, 1).astype(np.float64)
    y = band_target.flatten().astype(np.float64)
    xm, ym = X.mean(0), y.mean()
    coef, *_ = np.linalg.lstsq(X - xm, y - ym, rcond=None)
    recon = (X @ coef + (ym - xm @ coef)).reshape(shape)
    R = ratio_2c_match_c_from_sums_outlier(band_target, recon, zero_value_out=-.5)
    return np.where(band_target == 0.0, -.5, R)


def emit_rescale(magic, rgb):
    """emit_dataset.py:62-69,80-101: crop to x32, clip(mf/240,0,2)*1750, clip(rgb/20,0,2)*60, nan_to_num."""
    w, h = magic.shape
    w32, h32 = (w // 32) * 32, (h // 32) * 32
    e_magic = np.clip(magic[:w32, :h32] / 240., 0., 2.) * 1750.
    e_rgb = np.clip(rgb[:, :w32, :h32] / 20., 0., 2.) * 60.
    out = np.ones((4, w32, h32), dtype=np.float32)
    out[0] = e_magic
    out[1:] = e_rgb
    return np.nan_to_num(out)
