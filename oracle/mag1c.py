"""Restatement of ``starcop/models/mag1c.py``: ``rmf`` (:283-348), ``acrwl1mf`` (:176-280),
``func_by_groups`` (:116-174), ``get_mask_bad_bands`` (:98-113) and the feeding logic of
``starcop/process_aviris.py:183-219`` / ``starcop/models/mag1c_emit.py:40-84``.

Written with explicit torch-CPU linear algebra in the same operation order as the
reference (bmm / cholesky / cholesky_solve), dtype-generic (float32 for AVIRIS, float64 for
EMIT).  Test infrastructure only (see ``oracle/__init__.py``).
"""
import numpy as np
import torch

NODATA = -9999          # mag1c.py:55
SCALING = 1e5           # mag1c.py:56
EPSILON = 1e-9          # mag1c.py:57


def get_mask_bad_bands(wave):
    """mag1c.py:98-113."""
    wave = np.asarray(wave)
    return ~(((wave < 400) | (wave > 2485)) | (((wave > 1350) & (wave < 1420)) | ((wave > 1800) & (wave < 1945))))


def band_keep_aviris(wavelengths):
    """process_aviris.py:192-195: not water vapour, 2122..2488 nm; must be one contiguous slice (:203-206)."""
    wavelengths = np.asarray(wavelengths)
    keep = get_mask_bad_bands(wavelengths) & (wavelengths >= 2122) & (wavelengths <= 2488)   # bounds are kept (:194-195)
    idx = np.where(keep)[0]
    assert len(idx) and np.all(np.diff(idx) == 1), "band selection must be contiguous"
    return slice(int(idx[0]), int(idx[-1]) + 1)


def _cov_solve(modx, mu, target, N, alpha):
    d = modx - mu
    C = torch.bmm(d.transpose(1, 2), d) / N                                     # :245 / :317
    C = C.lerp_(torch.diag_embed(torch.diagonal(C, dim1=-2, dim2=-1)), alpha)   # :246 / :318
    L = torch.linalg.cholesky(C)
    return torch.cholesky_solve(target.transpose(1, 2), L)                      # [b,s,1]


@torch.no_grad()
def rmf(x, template, alpha=0., apply_scaling=True):
    """mag1c.py:283-348 with mask=None, albedo_override=False, zero_override=False."""
    N = x.shape[1]
    template = template[None, None]
    mu = torch.mean(x, 1, keepdim=True)
    target = template * mu
    Cit = _cov_solve(x.clone(), mu, target, N, alpha)
    normalizer = torch.bmm(target, Cit)
    R = torch.bmm(x, mu.transpose(1, 2)) / torch.bmm(mu, mu.transpose(1, 2))
    mf = torch.relu(torch.bmm(x - mu, Cit) / (R * normalizer))
    if apply_scaling:
        mf = mf * SCALING
    return mf, R


@torch.no_grad()
def acrwl1mf(x, template, num_iter=30, alpha=0., covariance_update_scaling=1.):
    """mag1c.py:176-280 with the default flags (mask=None, no overrides, no energy)."""
    N = x.shape[1]
    mf, R = rmf(x, template, alpha=alpha, apply_scaling=False)
    template = template[None, None]
    target = template * torch.mean(x, dim=1, keepdim=True)                      # :233
    for _ in range(num_iter):
        modx = x - covariance_update_scaling * R * mf * target                  # :241
        mu = torch.mean(modx, dim=1, keepdim=True)
        target = template * mu
        Cit = _cov_solve(modx, mu, target, N, alpha)
        regularizer = 1 / (R * (mf + EPSILON))                                  # :255
        normalizer = torch.bmm(target, Cit)
        if torch.sum(torch.lt(normalizer, 1)):                                  # :264-266
            normalizer = normalizer.clamp_(min=1)
        mf = torch.relu((torch.bmm(x - mu, Cit) - regularizer) / (R * normalizer))
    return mf * SCALING, R


@torch.no_grad()
def func_by_groups(func, x, groups, mask=None):
    """mag1c.py:116-174 (result-equivalent: the chunked bounding-box reads only bound I/O)."""
    groups = np.asarray(groups)
    albedo_out = torch.tensor(np.zeros(x.shape[:2], dtype=x.dtype) + NODATA)
    mf_out = albedo_out.clone()
    if mask is None:
        mask = np.all(x > NODATA, axis=-1)
    for g in np.sort(np.unique(groups[mask])):
        m = (groups == g) & mask
        if np.sum(m) <= 10:                                                     # :166
            continue
        mf_i, al_i = func(torch.tensor(x[m, :]).unsqueeze(0))
        mf_out[m] = mf_i[0, :, 0]
        albedo_out[m] = al_i[0, :, 0]
    return mf_out, albedo_out


def mag1c_tile_columns(cube_bip, template, band_slice, num_iter=30, alpha=0.):
    """A (H, W, C) BIP cube filtered with groups = image columns (detector columns,
    process_aviris.py:211-212 with an identity GLT): x[b=W, p=H, s]."""
    x = torch.as_tensor(np.ascontiguousarray(cube_bip[:, :, band_slice])).permute(1, 0, 2).contiguous()
    mf, R = acrwl1mf(x, torch.as_tensor(template, dtype=x.dtype), num_iter=num_iter, alpha=alpha)
    return mf[..., 0].T.contiguous(), R[..., 0].T.contiguous()                  # (H, W) each


@torch.no_grad()
def mag1c_emit(raw_data, wavelengths, template, fill_value_default=-9999.0, use_wavelength_range=(2122, 2488),
               num_iter=30, covariance_lerp_alpha=1e-4, column_step=None):
    """starcop/models/mag1c_emit.py:40-90 with georreferenced=False; raw_data: (rows, cols, bands) numpy;
    template: unit absorption spectrum of the SELECTED bands (the reference builds it at :45)."""
    wl = np.asarray(wavelengths)
    sel = (wl >= use_wavelength_range[0]) & (wl <= use_wavelength_range[1])                  # :40
    raw = np.asarray(raw_data)[..., sel]
    spec = torch.as_tensor(np.asarray(template, dtype=np.float64))
    invalid = np.any(raw == fill_value_default, axis=-1)                                     # :51
    mf_out = np.full(invalid.shape, dtype=np.float64, fill_value=fill_value_default)
    al_out = np.full(invalid.shape, dtype=np.float64, fill_value=fill_value_default)
    column_step = column_step or raw.shape[1]
    for c0 in range(0, raw.shape[1], column_step):                                           # :58-84
        c1 = min(c0 + column_step, raw.shape[1])
        valid = ~invalid[:, c0:c1]
        if not valid.any():
            continue
        x = torch.tensor(raw[:, c0:c1, :][valid, :][np.newaxis].astype(np.float64))
        mf, al = acrwl1mf(x, spec, num_iter=num_iter, alpha=covariance_lerp_alpha)
        mf_out[:, c0:c1][valid] = np.array(mf)[0, :, 0]
        al_out[:, c0:c1][valid] = np.array(al)[0, :, 0]
    return mf_out.astype(np.float32), al_out.astype(np.float32)
