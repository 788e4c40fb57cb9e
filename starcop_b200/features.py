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
