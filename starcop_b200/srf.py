"""Drop-in for ``starcop.data.aviris.transform_to_srf`` (aviris.py:262-338): simulate a multispectral sensor (WV3 /
Sentinel-2 bands) from a hyperspectral cube through spectral response functions.

Host part (numpy, one-off per sensor, :275-316): every SRF wavelength is assigned the NEAREST cube band
(``scipy.interpolate.interp1d(kind="nearest")``), responses <= 1e-4 are dropped, the rest are normalised to sum
one and summed per cube band -> a (K, C) weight table.  Device part (:320-326): out[k] = sum_c W[k,c] * cube[..., c]
with the reference's missing-value rule, ONE pass over the BIP cube (``sc_srf_aggregate``)."""
import numpy as np
import torch

from . import _lib


def srf_weight_table(srf_wavelengths, srf_responses, band_centers):
    """srf_wavelengths: (n,) nm (``srf.index``); srf_responses: (K, n) (the ``bands`` columns of the SRF frame);
    band_centers: (C,) nm of the cube's bands.  -> (K, C) float64 weights, rows sum to 1."""
    wl = np.asarray(srf_wavelengths, dtype=np.float64)
    resp = np.asarray(srf_responses, dtype=np.float64)
    centers = np.asarray(band_centers, dtype=np.float64)
    if wl.min() < centers.min() or wl.max() > centers.max():
        raise ValueError("A value in x_new is outside the interpolation range.")        # interp1d's bounds error
    # nearest band with scipy's tie rule (interp1d 'nearest' rounds half DOWN: x_new == midpoint -> the left band)
    order = np.argsort(centers)
    mids = (centers[order][1:] + centers[order][:-1]) / 2.0
    nearest = order[np.searchsorted(mids, wl, side="left")]
    K, C = resp.shape[0], centers.shape[0]
    W = np.zeros((K, C), dtype=np.float64)
    for k in range(K):
        keep = ~(resp[k] <= 1e-4)
        w = resp[k, keep] / resp[k, keep].sum()
        np.add.at(W[k], nearest[keep], w)
    return W


def transform_to_srf(cube_bip, weights, fill_value_default=0.0):
    """cube_bip: (..., H, W, C) CUDA float32 BIP radiance; weights: (K, C) table -> (..., K, H, W) float32."""
    if not cube_bip.is_cuda:
        raise _lib.StarcopB200Error("starcop_b200 runs on CUDA tensors only (no CPU path)")
    x = cube_bip.contiguous().float()
    *lead, H, W, C = x.shape
    wd, ranges, K = _device_table(weights, C, x.device)
    T = int(np.prod(lead)) if lead else 1
    out = torch.empty((T, K, H, W), dtype=torch.float32, device=x.device)
    # all stacked scenes in ONE launch; the output is planar (K, H, W) per scene, like the reference's GeoTensor
    _lib.call("sc_srf_aggregate", x.data_ptr(), T * H * W, H * W, C, wd.data_ptr(), ranges.ctypes.data, K, float(fill_value_default),
              out.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream)
    return out.view(*lead, K, H, W) if lead else out[0]


_TABLES = {}


def _device_table(weights, C, device):
    """(device f32 table, host band ranges, K) of a weight table, cached per table object / device"""
    key = (id(weights), str(device))
    ent = _TABLES.get(key)
    if ent is None or ent[0] is not weights:
        Wt = np.asarray(weights.detach().cpu() if torch.is_tensor(weights) else weights, dtype=np.float64)
        K = Wt.shape[0]
        assert Wt.shape[1] == C, f"weight table has {Wt.shape[1]} bands, cube has {C}"
        ranges = np.zeros((K, 2), dtype=np.int32)
        for k in range(K):
            nz = np.nonzero(Wt[k])[0]
            ranges[k] = (nz[0], nz[-1] + 1) if len(nz) else (0, 0)
        wd = torch.as_tensor(Wt.astype(np.float32), device=device).contiguous()
        if len(_TABLES) > 64:
            _TABLES.clear()
        ent = _TABLES[key] = (weights, wd, ranges, K)
    return ent[1], ent[2], ent[3]
