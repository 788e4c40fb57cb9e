"""Thin host-side helpers over the C ABI for callers that do not run inside the engine's arenas (tests, the
per-layer benchmarks): they size and allocate the workspaces an entry point asks for with torch, then call it."""
import ctypes

import torch

from . import _lib


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def tc_conv_wgrad(x, ldx, dy, lddy, dw, N, H, W, cin, cout, k, stride=1, workspace=None):
    """dW (OIHW f32) += X^T dY on the tensor cores (``sc_tc_conv_wgrad``): deterministic split-K.
    x / dy: bf16 NHWC tensors (channel strides ldx / lddy); ``workspace``: the partial-tile buffer to reuse."""
    lib = _lib.load()
    dev = dw.device
    if workspace is None:
        nb = lib.sc_tc_conv_wgrad_workspace_bytes(N, H, W, cin, cout, k, k, stride)
        if nb < 0:
            raise _lib.StarcopB200Error("sc_tc_conv_wgrad: unsupported shape")
        workspace = torch.empty(max(nb, 4), dtype=torch.uint8, device=dev)
    _lib.call("sc_tc_conv_wgrad", x.data_ptr(), ldx, dy.data_ptr(), lddy, dw.data_ptr(), workspace.data_ptr(),
              N, H, W, cin, cout, k, k, stride, _stream(dev))
    return workspace


def conv_wgrad(x, ldx, dy, lddy, dw, N, H, W, cin, cout, k, stride, pad, dtype, workspace=None):
    """fp32-FMA weight gradient (``sc_conv_wgrad``), deterministic pixel splits."""
    lib = _lib.load()
    dev = dw.device
    if workspace is None:
        nb = lib.sc_conv_wgrad_workspace_bytes(N, H, W, cin, cout, k, k, stride, pad)
        workspace = torch.empty(max(nb, 4), dtype=torch.uint8, device=dev)
    _lib.call("sc_conv_wgrad", x.data_ptr(), ldx, dy.data_ptr(), lddy, dw.data_ptr(), workspace.data_ptr(), N, H, W, cin, cout,
              k, k, stride, pad, dtype, _stream(dev))
    return workspace


def bce_loss_buffer(B, HW, device):
    """zero-filled loss accumulator of ``sc_bce_fused`` ([0] sum, [1] ticket, [2:] block partials)."""
    return torch.zeros(_lib.load().sc_bce_loss_words(B, HW), dtype=torch.float64, device=device)
