"""Drop-in for the hot functions of ``starcop/data/feature_extration.py`` on CUDA:
``weight_mag1c`` (:32-35) and ``ratio_2c_match_c_from_sums_outlier`` (:42-56), batched over tiles."""
import torch

from . import _lib


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def weight_mag1c(mag1c):
    """np.clip(mag1c / 400, 0.1, 1)."""
    if not mag1c.is_cuda:
        raise _lib.StarcopB200Error("starcop_b200 runs on CUDA tensors only (no CPU path)")
    m = mag1c.contiguous().float()
    out = torch.empty_like(m)
    _lib.call("sc_weight_mag1c", m.data_ptr(), out.data_ptr(), m.numel(), _stream(m.device))
    return out


def ratio_2c_match_c_from_sums_outlier(background_channel, signal, p=5, zero_value_out=-.6):
    """background_channel, signal: (..., H, W) CUDA float32 with identical shapes; every leading index
    is an independent tile (the reference is called once per tile).  Returns R of the same shape."""
    if not (background_channel.is_cuda and signal.is_cuda):
        raise _lib.StarcopB200Error("starcop_b200 runs on CUDA tensors only (no CPU path)")
    assert background_channel.shape == signal.shape
    bg = background_channel.contiguous().float()
    sg = signal.contiguous().float()
    H, W = bg.shape[-2:]
    T = bg.numel() // (H * W)
    out = torch.empty_like(bg)
    lib = _lib.load()
    ws = torch.empty(max(lib.sc_ratio_workspace_bytes(T, H * W), 8), dtype=torch.uint8, device=bg.device)
    _lib.call("sc_ratio_product", bg.data_ptr(), sg.data_ptr(), out.data_ptr(), T, H * W, float(p),
              float(zero_value_out), ws.data_ptr(), _stream(bg.device))
    return out


def ratio_MLR_local(bands_bg, band_target_signal):
    """feature_extration.py:58-124 with division="c_matched_outliers" (the registry's default):
    bands_bg: (..., K, H, W) or a list of K tensors (..., H, W); band_target_signal: (..., H, W)."""
    if isinstance(bands_bg, (list, tuple)):
        bands_bg = torch.stack(list(bands_bg), dim=-3)
    tgt = band_target_signal.contiguous().float()
    x = bands_bg.contiguous().float()
    if not (tgt.is_cuda and x.is_cuda):
        raise _lib.StarcopB200Error("starcop_b200 runs on CUDA tensors only (no CPU path)")
    H, W = tgt.shape[-2:]
    K = x.shape[-3]
    T = tgt.numel() // (H * W)
    lib = _lib.load()
    ws = torch.empty(lib.sc_mlr_workspace_bytes(T), dtype=torch.uint8, device=tgt.device)
    recon = torch.empty_like(tgt)
    st = _stream(tgt.device)
    _lib.call("sc_mlr_reconstruct", x.data_ptr(), tgt.data_ptr(), recon.data_ptr(), T, K, H * W, ws.data_ptr(), st)
    # ratio_2c_match_c_from_sums_outlier(band_target_signal, reconstruction, zero_value_out=-.5)  (:107-110)
    out = ratio_2c_match_c_from_sums_outlier(tgt, recon, zero_value_out=-.5)
    _lib.call("sc_zero_override", tgt.data_ptr(), out.data_ptr(), out.numel(), -0.5, st)
    return out


def ratio_MLR_local_5IN(IN1, IN2, IN3, IN4, IN5, target_B):
    return ratio_MLR_local([IN1, IN2, IN3, IN4, IN5], target_B)


def ratio_MLR_local_9IN(IN1, IN2, IN3, IN4, IN5, IN6, IN7, IN8, IN9, target_B):
    return ratio_MLR_local([IN1, IN2, IN3, IN4, IN5, IN6, IN7, IN8, IN9], target_B)


def emit_rescale(magic, rgb):
    """emit_tools/emit_dataset.py:62-101: magic (H,W), rgb (3,H,W) CUDA float32 -> (4, H32, W32) model input."""
    if not (magic.is_cuda and rgb.is_cuda):
        raise _lib.StarcopB200Error("starcop_b200 runs on CUDA tensors only (no CPU path)")
    m, r = magic.contiguous().float(), rgb.contiguous().float()
    H, W = m.shape
    out = torch.empty(4, H // 32 * 32, W // 32 * 32, dtype=torch.float32, device=m.device)
    _lib.call("sc_emit_rescale", m.data_ptr(), r.data_ptr(), out.data_ptr(), H, W, _stream(m.device))
    return out


# the reference's registry (feature_extration.py:193-246) for the products computed here
FEATURES = {
    "weight_mag1c": {"function": weight_mag1c, "inputs": ["mag1c"]},
    "ratio_aviris_2350_2310_out": {"function": ratio_2c_match_c_from_sums_outlier, "inputs": ["TOA_AVIRIS_2350nm", "TOA_AVIRIS_2310nm"]},
    "ratio_aviris_2350_2360_out": {"function": ratio_2c_match_c_from_sums_outlier, "inputs": ["TOA_AVIRIS_2350nm", "TOA_AVIRIS_2360nm"]},
    "ratio_aviris_2360_2310_out": {"function": ratio_2c_match_c_from_sums_outlier, "inputs": ["TOA_AVIRIS_2360nm", "TOA_AVIRIS_2310nm"]},
    "ratio_wv3_B7_B5_varon21_sum_c_out": {"function": ratio_2c_match_c_from_sums_outlier, "inputs": ["TOA_WV3_SWIR7", "TOA_WV3_SWIR5"]},
    "ratio_wv3_B8_B5_varon21_sum_c_out": {"function": ratio_2c_match_c_from_sums_outlier, "inputs": ["TOA_WV3_SWIR8", "TOA_WV3_SWIR5"]},
    "ratio_wv3_B7_B6_varon21_sum_c_out": {"function": ratio_2c_match_c_from_sums_outlier, "inputs": ["TOA_WV3_SWIR7", "TOA_WV3_SWIR6"]},
    "ratio_wv3_B7_B7MLR_SanchezGarcia22_sum_c_out": {"function": ratio_MLR_local_5IN, "inputs": ["TOA_WV3_SWIR1", "TOA_WV3_SWIR2", "TOA_WV3_SWIR4", "TOA_WV3_SWIR5", "TOA_WV3_SWIR6", "TOA_WV3_SWIR7"]},
    "ratio_wv3_B8_B8MLR_SanchezGarcia22_sum_c_out": {"function": ratio_MLR_local_5IN, "inputs": ["TOA_WV3_SWIR1", "TOA_WV3_SWIR2", "TOA_WV3_SWIR4", "TOA_WV3_SWIR5", "TOA_WV3_SWIR6", "TOA_WV3_SWIR8"]},
}
