"""Drop-in for the hot functions of ``starcop/models/mag1c.py`` on the CUDA matched-filter kernel
(``sc_mag1c_filter``): ``acrwl1mf`` (:176-280), ``rmf`` (:283-348), ``func_by_groups`` (:116-174),
``get_mask_bad_bands`` (:98-113), plus the tile driver used by ``run_mag1c``
(starcop/process_aviris.py:183-219: BIP cube, contiguous band slice, groups = detector columns).

Template generation (``generate_template_from_bands``, :60-95, a one-off computation over the
31800-sample CH4 look-up table) is host numpy here as in the reference; the filter takes its result
as an array.
"""
import os
import re

import numpy as np
import torch

from . import _lib

NODATA = -9999
SCALING = 1e5
EPSILON = 1e-9


CH4_CONCENTRATIONS = (0, 500, 1000, 2000, 4000, 8000, 16000)      # ppm*m columns of ch4.lut (mag1c.py:79)


def read_ch4_lut(lut_dir=None):
    """The reference ships its radiative-transfer look-up table as an ENVI pair ``ch4.hdr`` / ``ch4.lut``
    next to ``starcop/models/mag1c.py`` and reads it with ``spectral`` (:75-78).  Same data without that
    dependency: little-endian float64, BIP with one sample per line -> (7 concentrations, n wavelengths),
    wavelengths (nm) from the header's ``wavelength = {...}`` list.
    ``lut_dir`` defaults to $STARCOP_CH4_LUT_DIR, then to an importable ``starcop.models`` package."""
    if lut_dir is None:
        lut_dir = os.environ.get("STARCOP_CH4_LUT_DIR")
    if lut_dir is None:
        try:
            import starcop.models as _m          # the reference package, if it is installed
            lut_dir = os.path.dirname(os.path.abspath(_m.__file__))
        except Exception as e:                   # noqa: BLE001
            raise FileNotFoundError("ch4.hdr / ch4.lut not found: pass lut_dir= or set STARCOP_CH4_LUT_DIR") from e
    with open(os.path.join(lut_dir, "ch4.hdr")) as f:
        hdr = f.read()
    m = re.search(r"wavelength\s*=\s*\{([^}]*)\}", hdr)
    if m is None:
        raise ValueError("ch4.hdr has no wavelength list")
    wave = np.array([float(v) for v in m.group(1).replace("\n", " ").split(",") if v.strip()], dtype=np.float64)
    raw = np.fromfile(os.path.join(lut_dir, "ch4.lut"), dtype="<f8")
    nconc = len(CH4_CONCENTRATIONS)
    if raw.size != wave.size * nconc:
        raise ValueError(f"ch4.lut holds {raw.size} values, expected {wave.size} wavelengths x {nconc} concentrations")
    return raw.reshape(wave.size, nconc).T.copy(), wave


def generate_template_from_bands(centers, fwhm, lut=None, lut_dir=None):
    """mag1c.py:60-95: the CH4 unit absorption spectrum of a sensor's bands.  Each band is a Gaussian
    spectral response (sigma = fwhm / 2.355) normalised to unit sum over the LUT's wavelength grid; the
    LUT radiances are resampled through it, and the per-band slope of log(radiance) against concentration
    (least squares over the 7 LUT columns) times 1e5 is the template.  Returns (K, 2): [center, spectrum].
    ``lut`` = (rads (7, n), wave (n,)) overrides the files."""
    centers = np.asarray(centers)
    fwhm = np.asarray(fwhm)
    if np.any(~np.isfinite(centers)) or np.any(~np.isfinite(fwhm)):
        raise RuntimeError("Band Wavelengths Centers/FWHM data contains non-finite data (NaN or Inf).")
    if centers.shape[0] != fwhm.shape[0]:
        raise RuntimeError("Length of band center wavelengths and band fwhm arrays must be equal.")
    rads, wave = lut if lut is not None else read_ch4_lut(lut_dir)
    conc = np.asarray(CH4_CONCENTRATIONS)
    var = (fwhm / (2.0 * np.sqrt(2.0 * np.log(2.0)))) ** 2
    resp = np.exp(-(wave[:, None] - centers[None, :]) ** 2 / (2 * var)) / (2 * np.pi * var) ** 0.5
    tot = resp.sum(axis=0)
    resp = np.divide(resp, tot, where=tot > 0)
    resampled = np.asarray(rads).dot(resp)
    lograd = np.log(resampled, where=resampled > 0)
    design = np.stack((np.ones_like(conc), conc)).T
    slope, _, _, _ = np.linalg.lstsq(design, lograd, rcond=None)
    return np.stack((centers, slope[1, :] * SCALING)).T


def get_mask_bad_bands(wave):
    """mag1c.py:98-113: keep 400..2485 nm minus the 1350-1420 / 1800-1945 nm water bands."""
    wave = np.asarray(wave)
    return ~(((wave < 400) | (wave > 2485)) | (((wave > 1350) & (wave < 1420)) | ((wave > 1800) & (wave < 1945))))


def band_keep_aviris(wavelengths):
    """process_aviris.py:192-206: the matched-filter window as ONE contiguous slice."""
    wavelengths = np.asarray(wavelengths)
    # band_keep[wavelengths < lo] = False; band_keep[wavelengths > hi] = False  -> the bounds themselves are kept
    keep = get_mask_bad_bands(wavelengths) & (wavelengths >= 2122) & (wavelengths <= 2488)
    idx = np.where(keep)[0]
    if not len(idx) or np.any(np.diff(idx) != 1):
        raise AssertionError("Selected bands are not contiguous")
    return slice(int(idx[0]), int(idx[-1]) + 1)


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _filter(x_flat, pixel_stride, pix_idx, counts, template, mf_out, al_out, S, num_iter, alpha, skip_le=10):
    dev = x_flat.device
    fp64 = x_flat.dtype == torch.float64
    tmpl = torch.as_tensor(np.asarray(template, dtype=np.float64) if not torch.is_tensor(template)
                           else template.detach().double().cpu().numpy()).to(dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    G, pmax = pix_idx.shape
    _lib.call("sc_mag1c_filter", x_flat.data_ptr(), pixel_stride, pix_idx.data_ptr(),
              counts.data_ptr() if counts is not None else 0, pmax, tmpl.data_ptr(), mf_out.data_ptr(),
              al_out.data_ptr(), G, S, num_iter, float(alpha), int(fp64), int(skip_le), status.data_ptr(), _stream(dev))
    return status


@torch.no_grad()
def acrwl1mf(x, template, num_iter=30, alpha=0., check=True):
    """x: [b, p, s] CUDA tensor (float32 or float64) -> (mf [b,p,1] in ppm*m, albedo R [b,p,1])."""
    if not x.is_cuda:
        raise _lib.StarcopB200Error("starcop_b200 runs on CUDA tensors only (no CPU path)")
    assert x.dim() == 3, "x must be [batch(groups), pixels, spectrum]"
    x = x.contiguous()
    b, p, s = x.shape
    if p <= 10:
        # the kernel skips groups of <= 10 pixels (func_by_groups' rule, mag1c.py:166-168); a DIRECT call with so few
        # pixels has a rank-deficient covariance (p < s): the reference raises from torch.linalg.cholesky there
        raise torch.linalg.LinAlgError(f"linalg.cholesky: covariance of {p} pixels x {s} bands is not positive-definite")
    idx = torch.arange(b * p, dtype=torch.int32, device=x.device).view(b, p)
    mf = torch.empty(b, p, 1, dtype=x.dtype, device=x.device)
    al = torch.empty_like(mf)
    status = _filter(x, s, idx, None, template, mf, al, s, num_iter, alpha)
    if check and int(status.item()):
        # the reference raises from torch.linalg.cholesky in the same situation
        raise torch.linalg.LinAlgError(f"linalg.cholesky: covariance of {int(status.item())} group(s) is not positive-definite")
    return mf, al


@torch.no_grad()
def rmf(x, template, alpha=0., check=True):
    """mag1c.py:283-348 with the default flags: the plain (albedo-normalised) matched filter."""
    return acrwl1mf(x, template, num_iter=0, alpha=alpha, check=check)


@torch.no_grad()
def func_by_groups(x, groups, template, mask=None, num_iter=30, alpha=0.):
    """mag1c.py:116-174 for func = acrwl1mf: x (H, W, S) CUDA radiance, groups (H, W) integer map,
    mask (H, W) valid pixels -> (mf, albedo), each (H, W), NODATA where not computed.
    The gather lists are built on the host like the reference's loops; ALL groups then run in one launch."""
    H, W, S = x.shape
    dev = x.device
    g = np.asarray(groups.cpu() if torch.is_tensor(groups) else groups)
    if mask is None:
        m = torch.all(x > NODATA, dim=-1).cpu().numpy()
    else:
        m = np.asarray(mask.cpu() if torch.is_tensor(mask) else mask).astype(bool)
    flat = np.arange(H * W, dtype=np.int32).reshape(H, W)
    ids = np.sort(np.unique(g[m]))
    lists = [flat[(g == i) & m] for i in ids]
    pmax = max((len(l) for l in lists), default=1)
    idx = np.zeros((max(len(lists), 1), pmax), dtype=np.int32)
    cnt = np.zeros(max(len(lists), 1), dtype=np.int32)
    for k, l in enumerate(lists):
        idx[k, :len(l)] = l
        cnt[k] = len(l)
    mf = torch.full((H, W), float(NODATA), dtype=x.dtype, device=dev)
    al = torch.full((H, W), float(NODATA), dtype=x.dtype, device=dev)
    if len(lists):
        xc = x.contiguous()
        _filter(xc, S, torch.from_numpy(idx).to(dev), torch.from_numpy(cnt).to(dev), template, mf, al, S, num_iter, alpha)
    return mf, al


_IDX_CACHE = {}


def _column_groups(n, H, W, dev):
    """pixel lists for groups = image columns of n stacked (H, W) tiles: group (t, w) = pixels (t, :, w)."""
    key = (n, H, W, str(dev))
    if key not in _IDX_CACHE:
        t = torch.arange(n, device=dev, dtype=torch.int64)[:, None, None] * (H * W)
        w = torch.arange(W, device=dev, dtype=torch.int64)[None, :, None]
        h = torch.arange(H, device=dev, dtype=torch.int64)[None, None, :] * W
        _IDX_CACHE[key] = (t + w + h).to(torch.int32).reshape(n * W, H).contiguous()
    return _IDX_CACHE[key]


@torch.no_grad()
def mag1c_tiles(cube, template, band_slice, num_iter=30, alpha=0.):
    """cube: (n, H, W, C) BIP radiance on the GPU (process_aviris.py:183-184); the filter runs on the
    contiguous window ``band_slice`` with one group per image column (identity GLT), straight from
    the cube's memory (no band gather, no transpose).  -> (mf, albedo), each (n, H, W)."""
    if not cube.is_cuda:
        raise _lib.StarcopB200Error("starcop_b200 runs on CUDA tensors only (no CPU path)")
    cube = cube.contiguous()
    n, H, W, C = cube.shape
    S = band_slice.stop - band_slice.start
    idx = _column_groups(n, H, W, cube.device)
    # groups (columns) of <= 10 pixels are skipped like func_by_groups does: their outputs keep the NODATA fill
    mk = torch.full if H <= 10 else (lambda shape, _v, **kw: torch.empty(shape, **kw))
    mf = mk((n, H, W), float(NODATA), dtype=cube.dtype, device=cube.device)
    al = mk((n, H, W), float(NODATA), dtype=cube.dtype, device=cube.device)
    xw = cube.view(-1)[band_slice.start:]          # same storage, offset to the first window band
    _filter(xw, C, idx, None, template, mf, al, S, num_iter, alpha)
    return mf, al


DEFAULT_WAVELENGTH_RANGE = (2122, 2488)


@torch.no_grad()
def mag1c_emit(raw_data, wavelengths, template=None, fwhm=None, fill_value_default=-9999.0,
               use_wavelength_range=DEFAULT_WAVELENGTH_RANGE, num_iter=30, covariance_lerp_alpha=1e-4, column_step=None,
               lut_dir=None, check=True):
    """``starcop.models.mag1c_emit.mag1c_emit`` (mag1c_emit.py:16-90) on a raw (rows, cols, bands) EMIT cube that is
    already on the GPU (the georeader ``EMITImage`` container and the geo-referencing of the result are out of
    scope: this is the ``georreferenced=False`` path the inference notebook uses).

    Bands with ``lo <= wavelength <= hi`` are selected (:40), pixels with any selected band equal to the fill value
    are invalid (:51), the cube is processed in float64 (:75) in groups of ``column_step`` detector columns (:56-84;
    every group with at least one valid pixel is filtered: there is no <= 10 pixel rule here), with diagonal
    loading ``alpha = 1e-4`` (:18).  ``template``: the unit absorption spectrum of the SELECTED bands, or None to
    build it from ``fwhm`` with ``generate_template_from_bands`` (:45).  Returns (mf, albedo), each (rows, cols)
    float32 with the fill value where nothing was computed (:90)."""
    if not raw_data.is_cuda:
        raise _lib.StarcopB200Error("starcop_b200 runs on CUDA tensors only (no CPU path)")
    wl = np.asarray(wavelengths, dtype=np.float64)
    sel = (wl >= use_wavelength_range[0]) & (wl <= use_wavelength_range[1])
    assert sel.any(), "There are no bands in the selected wavelength range"
    idx = np.where(sel)[0]
    if template is None:
        assert fwhm is not None, "pass the template of the selected bands, or the bands' fwhm to build it"
        template = generate_template_from_bands(wl[sel], np.asarray(fwhm, dtype=np.float64)[sel], lut_dir=lut_dir)[:, 1]
    template = np.asarray(template, dtype=np.float64)
    S = len(idx)
    assert template.shape[0] == S, f"template has {template.shape[0]} bands, {S} selected"
    rows, cols, _ = raw_data.shape
    dev = raw_data.device
    contiguous = bool(np.all(np.diff(idx) == 1))
    x = (raw_data[..., int(idx[0]):int(idx[-1]) + 1] if contiguous else raw_data[..., torch.as_tensor(idx, device=dev)])
    invalid = torch.any(x == fill_value_default, dim=-1)                                       # (rows, cols)
    x64 = x.to(torch.float64).contiguous()                                                     # :75
    column_step = column_step or cols
    G, pmax = (cols + column_step - 1) // column_step, rows * column_step
    # per group: the valid pixels of raw[:, c0:c1] in row-major order of the slice (``raw_data_slice[valid_slice]``,
    # mag1c_emit.py:56-84).  Built for ALL groups at once on the device (the reference loops over the groups on the
    # host; a per-group loop here would cost one device synchronisation per group -- 621 for a granule): the columns
    # are padded to G * column_step with invalid pixels, viewed as (G, rows * column_step), and each row is
    # compacted by a stable sort of the validity flags.
    cpad = G * column_step
    valid = torch.zeros(rows, cpad, dtype=torch.bool, device=dev)
    valid[:, :cols] = ~invalid
    flat = torch.arange(rows, device=dev, dtype=torch.int32)[:, None] * cols + torch.arange(cpad, device=dev, dtype=torch.int32)[None, :]
    vg = valid.view(rows, G, column_step).permute(1, 0, 2).reshape(G, pmax)
    fg = flat.view(rows, G, column_step).permute(1, 0, 2).reshape(G, pmax)
    order = torch.argsort((~vg).to(torch.uint8), dim=1, stable=True)            # valid pixels first, order kept
    pix = torch.gather(fg, 1, order).contiguous()
    cnt = vg.sum(dim=1).to(torch.int32)
    mf = torch.full((rows, cols), float(fill_value_default), dtype=torch.float64, device=dev)
    al = torch.full((rows, cols), float(fill_value_default), dtype=torch.float64, device=dev)
    status = _filter(x64, S, pix, cnt, template, mf, al, S, num_iter, covariance_lerp_alpha, skip_le=0)
    if check and int(status.item()):
        raise torch.linalg.LinAlgError(f"linalg.cholesky: covariance of {int(status.item())} group(s) is not positive-definite")
    return mf.float(), al.float()
