"""tcgen05 tensor-core convolution kernels (bf16 in, fp32 accumulate in TMEM) against PyTorch fp32
on the same bf16-rounded operands."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from starcop_b200 import ops  # noqa: E402
from starcop_b200._lib import call, load  # noqa: E402

DEV = "cuda"


def st():
    return torch.cuda.current_stream().cuda_stream


CASES = [  # N, H, W, Cin, Cout, k
    (1, 8, 16, 64, 64, 3),        # SW128, single tile
    (2, 16, 32, 32, 32, 3),       # SW64
    (2, 16, 16, 16, 16, 3),       # SW32 (decoder block 4 conv2 shape class)
    (1, 8, 16, 96, 48, 3),        # Cin % 64 != 0 -> 32-channel chunks; N tile 48
    (2, 8, 16, 128, 256, 1),      # pointwise
    (1, 8, 16, 1376, 256, 3),     # decoder block 0 conv1 channel geometry
    (1, 8, 16, 256, 1376, 3),     # its dgrad: 6 N tiles, ragged last tile
    (3, 32, 32, 80, 32, 3),       # many tiles, 2 accumulator stages, persistent loop
    (2, 16, 16, 320, 1280, 1),    # features.18
    (2, 16, 16, 152, 64, 3),      # decoder block 2 conv1: Cin % 16 != 0 (last chunk zero-filled by TMA)
    (2, 16, 16, 96, 24, 1),       # Cout % 16 != 0 (masked half tile)
    (2, 16, 16, 24, 144, 1),
]


def make(N, H, W, Cin, Cout, k, seed=0):
    torch.manual_seed(seed)
    x = torch.randn(N, H, W, Cin, device=DEV).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, k, k, device=DEV) / np.sqrt(Cin * k * k))
    return x, w


@pytest.mark.parametrize("N,H,W,Cin,Cout,k", CASES)
@pytest.mark.parametrize("with_stats", [False, True])
def test_tc_fprop(N, H, W, Cin, Cout, k, with_stats):
    assert load().sc_tc_supported() == 1
    x, w = make(N, H, W, Cin, Cout, k)
    cpad = load().sc_tc_cin_pad(Cin)
    wb = torch.empty(Cout * k * k * cpad, device=DEV, dtype=torch.bfloat16)
    call("sc_tc_pack_weights", w.data_ptr(), wb.data_ptr(), Cout, Cin, k, k, 0, cpad, Cout, st())
    wq = wb.view(Cout, k, k, cpad).permute(0, 3, 1, 2).float()[:, :Cin].contiguous()
    assert torch.equal(wq, w.to(torch.bfloat16).float())
    y = torch.full((N, H, W, Cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    import ctypes
    part = torch.full((load().sc_bn_partials_bytes(Cout) // 8,), float("nan"), dtype=torch.float64, device=DEV)
    nrows = ctypes.c_int(0)
    call("sc_tc_conv_fprop", x.data_ptr(), Cin, wb.data_ptr(), y.data_ptr(), Cout, part.data_ptr() if with_stats else 0,
         ctypes.byref(nrows), N, H, W, Cin, Cout, k, k, 1, 0, st())
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wq, padding=k // 2).permute(0, 2, 3, 1)
    err = (y.float() - ref).abs().max().item()
    assert err <= 2e-2 * max(1.0, ref.abs().max().item()), err
    # bf16 rounding of an fp32-accumulated result: at most 1 bf16 ulp off the rounded reference
    assert torch.allclose(y.float(), ref.to(torch.bfloat16).float(), rtol=1.6e-2, atol=1e-3)
    if with_stats:
        yf = y.double().reshape(-1, Cout)
        stats = part[:nrows.value * 2 * Cout].view(nrows.value, 2 * Cout).sum(0)
        assert torch.allclose(stats[:Cout], yf.sum(0), rtol=1e-4, atol=1e-2)
        assert torch.allclose(stats[Cout:], (yf * yf).sum(0), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("Cin,Cout", [(64, 32), (152, 64), (24, 96)])
def test_tc_dgrad_via_flipped_weights(Cin, Cout):
    N, H, W, k = 2, 16, 16, 3
    x, w = make(N, H, W, Cin, Cout, k)
    dy = torch.randn(N, H, W, Cout, device=DEV).to(torch.bfloat16)
    cpad = load().sc_tc_cin_pad(Cout)
    wt = torch.empty(Cin * k * k * cpad, device=DEV, dtype=torch.bfloat16)
    call("sc_tc_pack_weights", w.data_ptr(), wt.data_ptr(), Cout, Cin, k, k, 1, Cin, cpad, st())
    dx = torch.empty(N, H, W, Cin, device=DEV, dtype=torch.bfloat16)
    call("sc_tc_conv_fprop", dy.data_ptr(), Cout, wt.data_ptr(), dx.data_ptr(), Cin, 0, 0, N, H, W, Cout, Cin, k, k, 1, 0, st())
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.conv2d(xr, w.to(torch.bfloat16).float(), padding=1)
    (gx,) = torch.autograd.grad(yr, xr, dy.float().permute(0, 3, 1, 2))
    assert torch.allclose(dx.float(), gx.permute(0, 2, 3, 1), rtol=2e-2, atol=2e-2)
    # accumulate: dx += dgrad
    call("sc_tc_conv_fprop", dy.data_ptr(), Cout, wt.data_ptr(), dx.data_ptr(), Cin, 0, 0, N, H, W, Cout, Cin, k, k, 1, 1, st())
    assert torch.allclose(dx.float(), 2 * gx.permute(0, 2, 3, 1), rtol=3e-2, atol=4e-2)


# the last three: many pixel splits per gradient block -> the grouped split reduction with its last-arrival tail
@pytest.mark.parametrize("N,H,W,Cin,Cout,k", CASES + [(16, 64, 64, 32, 16, 3), (4, 32, 32, 16, 16, 3), (8, 128, 128, 24, 144, 1),
                                                 (4, 128, 128, 64, 64, 3), (4, 256, 256, 16, 96, 1)])
def test_tc_wgrad(N, H, W, Cin, Cout, k):
    x, w = make(N, H, W, Cin, Cout, k, seed=1)
    dy = torch.randn(N, H, W, Cout, device=DEV).to(torch.bfloat16)
    dw = torch.zeros(Cout, Cin, k, k, device=DEV)
    ws = ops.tc_conv_wgrad(x, Cin, dy, Cout, dw, N, H, W, Cin, Cout, k)
    torch.cuda.synchronize()
    wr = w.clone().requires_grad_(True)
    yr = F.conv2d(x.float().permute(0, 3, 1, 2), wr, padding=k // 2)
    (gw,) = torch.autograd.grad(yr, wr, dy.float().permute(0, 3, 1, 2))
    err = (dw - gw).abs().max().item()
    assert err <= 2e-3 * gw.abs().max().item(), (err, gw.abs().max().item())
    # deterministic split-K: a second launch (same workspace: the tickets must be back at zero) is bit-identical,
    # and the gradient is accumulated onto what is there
    dw2 = torch.zeros_like(dw)
    ops.tc_conv_wgrad(x, Cin, dy, Cout, dw2, N, H, W, Cin, Cout, k, workspace=ws)
    assert torch.equal(dw, dw2)
    ops.tc_conv_wgrad(x, Cin, dy, Cout, dw2, N, H, W, Cin, Cout, k, workspace=ws)
    assert torch.allclose(dw2, 2 * dw, rtol=1e-6, atol=0)


def test_tc_stem_stride2_fprop_and_wgrad():
    """encoder stem: Conv2d(4, 32, 3, stride 2, pad 1) on an input stored with 8-channel stride."""
    N, H, W, Cin, Cout, k = 2, 32, 64, 4, 32, 3
    torch.manual_seed(3)
    x8 = torch.zeros(N, H, W, 8, device=DEV, dtype=torch.bfloat16)
    x8[..., :Cin] = torch.randn(N, H, W, Cin, device=DEV).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, k, k, device=DEV) / 6
    cpad = load().sc_tc_cin_pad(Cin)
    wb = torch.empty(Cout * k * k * cpad, device=DEV, dtype=torch.bfloat16)
    call("sc_tc_pack_weights", w.data_ptr(), wb.data_ptr(), Cout, Cin, k, k, 0, cpad, Cout, st())
    y = torch.full((N, H // 2, W // 2, Cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    call("sc_tc_conv_fprop", x8.data_ptr(), 8, wb.data_ptr(), y.data_ptr(), Cout, 0, 0, N, H, W, Cin, Cout, k, k, 2, 0, st())
    xr = x8[..., :Cin].float().permute(0, 3, 1, 2)
    wr = w.to(torch.bfloat16).float().requires_grad_(True)
    ref = F.conv2d(xr, wr, stride=2, padding=1)
    assert torch.allclose(y.float(), ref.permute(0, 2, 3, 1), rtol=2e-2, atol=2e-2)
    dy = torch.randn(N, H // 2, W // 2, Cout, device=DEV).to(torch.bfloat16)
    dw = torch.zeros(Cout, Cin, k, k, device=DEV)
    ops.tc_conv_wgrad(x8, 8, dy, Cout, dw, N, H, W, Cin, Cout, k, stride=2)
    (gw,) = torch.autograd.grad(ref, wr, dy.float().permute(0, 3, 1, 2))
    assert (dw - gw).abs().max().item() <= 2e-3 * gw.abs().max().item()


@pytest.mark.parametrize("N,H,W,Cin", [(2, 32, 64, 4), (1, 64, 32, 3), (2, 64, 96, 1)])
def test_stem_space_to_depth_equals_the_stride2_convolution(N, H, W, Cin):
    """the stem as a space-to-depth convolution (sc_stem_s2d + equivalent 3x3 weights over 16 channels on the halo
    kernels): same outputs and the same weight gradient as Conv2d(Cin, 32, 3, stride 2, pad 1)"""
    Cout = 32
    torch.manual_seed(5)
    x8 = torch.zeros(N, H, W, 8, device=DEV, dtype=torch.bfloat16)
    x8[..., :Cin] = torch.randn(N, H, W, Cin, device=DEV).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV) / 6
    H2, W2 = H // 2, W // 2
    xs = torch.full((N, H2, W2, 16), float("nan"), device=DEV, dtype=torch.bfloat16)
    call("sc_stem_s2d", x8.data_ptr(), 8, Cin, xs.data_ptr(), N, H, W, st())
    wb = torch.empty(Cout * 144, device=DEV, dtype=torch.bfloat16)
    call("sc_stem_s2d_pack_weights", w.data_ptr(), wb.data_ptr(), Cout, Cin, st())
    y = torch.full((N, H2, W2, Cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    import ctypes
    n = ctypes.c_int(0)
    call("sc_tc_conv3x3_halo", xs.data_ptr(), 16, wb.data_ptr(), y.data_ptr(), Cout, 0, ctypes.byref(n), N, H2, W2, 16, Cout, 0, st())
    xr = x8[..., :Cin].float().permute(0, 3, 1, 2)
    wr = w.to(torch.bfloat16).float().requires_grad_(True)
    ref = F.conv2d(xr, wr, stride=2, padding=1)
    assert torch.allclose(y.float(), ref.permute(0, 2, 3, 1), rtol=2e-2, atol=2e-2)
    dy = torch.randn(N, H2, W2, Cout, device=DEV).to(torch.bfloat16)
    g16 = torch.zeros(Cout, 16, 3, 3, device=DEV)
    ops.tc_conv_wgrad(xs, 16, dy, Cout, g16, N, H2, W2, 16, Cout, 3)
    dw = torch.zeros(Cout, Cin, 3, 3, device=DEV)
    call("sc_stem_s2d_unpack_grad", g16.data_ptr(), dw.data_ptr(), Cout, Cin, st())
    (gw,) = torch.autograd.grad(ref, wr, dy.float().permute(0, 3, 1, 2))
    assert (dw - gw).abs().max().item() <= 2e-3 * gw.abs().max().item()
    # the positions of the equivalent weights that carry no original tap have zero weight (their gradient is ignored)
    wd = wb.float().view(Cout, 9, 16)
    assert int((wd != 0).sum()) <= Cout * 9 * Cin


HALO_CASES = [  # N, H, W, Cin, Cout
    (1, 16, 8, 16, 16),       # one tile, one K pair
    (2, 32, 32, 32, 16),      # decoder block 4 conv1 geometry, several tiles
    (2, 32, 24, 16, 32),      # its dgrad geometry
    (3, 48, 40, 80, 32),      # decoder block 3 conv1: 5 K pairs, persistent loop over > 2 accumulator stages
    (1, 20, 12, 24, 48),      # ragged tile edges (H % 16, W % 8 != 0) and a zero-filled half K pair (Cin % 16 != 0)
    (2, 16, 16, 64, 64),      # decoder block 2 conv2: 72 KB of resident weights
]


@pytest.mark.parametrize("N,H,W,Cin,Cout", HALO_CASES)
@pytest.mark.parametrize("with_stats", [False, True])
def test_tc_halo_fprop(N, H, W, Cin, Cout, with_stats):
    """thin-layer 3x3 kernel (one halo patch per tile, resident weights) against PyTorch on the same bf16 operands"""
    import ctypes
    lib = load()
    assert lib.sc_tc_halo_supported(Cin, Cout) == 1
    x, w = make(N, H, W, Cin, Cout, 3)
    cpad = lib.sc_tc_halo_cin_pad(Cin)
    wb = torch.empty(Cout * 9 * cpad, device=DEV, dtype=torch.bfloat16)
    call("sc_tc_pack_weights", w.data_ptr(), wb.data_ptr(), Cout, Cin, 3, 3, 0, cpad, Cout, st())
    y = torch.full((N, H, W, Cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    part = torch.full((lib.sc_bn_partials_bytes(Cout) // 8,), float("nan"), dtype=torch.float64, device=DEV)
    nrows = ctypes.c_int(0)
    call("sc_tc_conv3x3_halo", x.data_ptr(), Cin, wb.data_ptr(), y.data_ptr(), Cout, part.data_ptr() if with_stats else 0,
         ctypes.byref(nrows), N, H, W, Cin, Cout, 0, st())
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1)
    assert torch.allclose(y.float(), ref.to(torch.bfloat16).float(), rtol=1.6e-2, atol=1e-3)
    if with_stats:
        yf = y.double().reshape(-1, Cout)
        stats = part[:nrows.value * 2 * Cout].view(nrows.value, 2 * Cout).sum(0)
        assert torch.allclose(stats[:Cout], yf.sum(0), rtol=1e-4, atol=1e-2)
        assert torch.allclose(stats[Cout:], (yf * yf).sum(0), rtol=1e-4, atol=1e-2)
    # accumulate flag (gradient fan-in) and a channel-strided input / output (slices of a concat buffer)
    xw = torch.zeros(N, H, W, Cin + 8, device=DEV, dtype=torch.bfloat16)
    xw[..., 8:] = x
    yw = torch.zeros(N, H, W, Cout + 16, device=DEV, dtype=torch.bfloat16)
    yw[..., :Cout] = y
    call("sc_tc_conv3x3_halo", xw.data_ptr() + 16, Cin + 8, wb.data_ptr(), yw.data_ptr(), Cout + 16, 0, 0, N, H, W, Cin, Cout,
         1, st())
    assert torch.allclose(yw[..., :Cout].float(), 2 * ref, rtol=3e-2, atol=4e-2)
    assert float(yw[..., Cout:].abs().max()) == 0.0


def test_tc_halo_dgrad_via_flipped_weights():
    N, H, W, Cin, Cout = 2, 32, 32, 32, 16
    lib = load()
    x, w = make(N, H, W, Cin, Cout, 3)
    dy = torch.randn(N, H, W, Cout, device=DEV).to(torch.bfloat16)
    cpad = lib.sc_tc_halo_cin_pad(Cout)
    wt = torch.empty(Cin * 9 * cpad, device=DEV, dtype=torch.bfloat16)
    call("sc_tc_pack_weights", w.data_ptr(), wt.data_ptr(), Cout, Cin, 3, 3, 1, Cin, cpad, st())
    dx = torch.empty(N, H, W, Cin, device=DEV, dtype=torch.bfloat16)
    call("sc_tc_conv3x3_halo", dy.data_ptr(), Cout, wt.data_ptr(), dx.data_ptr(), Cin, 0, 0, N, H, W, Cout, Cin, 0, st())
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.conv2d(xr, w.to(torch.bfloat16).float(), padding=1)
    (gx,) = torch.autograd.grad(yr, xr, dy.float().permute(0, 3, 1, 2))
    assert torch.allclose(dx.float(), gx.permute(0, 2, 3, 1), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 16, 32, 16, 16), (2, 32, 32, 32, 16), (2, 16, 16, 32, 32), (1, 24, 48, 80, 32),
                                            (3, 8, 16, 16, 32), (2, 40, 64, 24, 16), (16, 64, 64, 32, 16)])
def test_tc_wgrad_halo(N, H, W, Cin, Cout, monkeypatch):
    """thin-layer weight gradient: halo patch fetched once, horizontal taps stacked along the MMA's M dimension.
    Against PyTorch fp32 on the same bf16 operands, against the per-tap tcgen05 kernel, and bit-reproducible."""
    lib = load()
    assert lib.sc_tc_wgrad_halo_supported(N, H, W, Cin, Cout) == 1
    torch.manual_seed(11)
    x = torch.randn(N, H, W, Cin, device=DEV).to(torch.bfloat16)
    dy = torch.randn(N, H, W, Cout, device=DEV).to(torch.bfloat16)
    dw = torch.zeros(Cout, Cin, 3, 3, device=DEV)
    ws = ops.tc_conv_wgrad(x, Cin, dy, Cout, dw, N, H, W, Cin, Cout, 3)
    wr = torch.zeros(Cout, Cin, 3, 3, device=DEV, requires_grad=True)
    yr = F.conv2d(x.float().permute(0, 3, 1, 2), wr, padding=1)
    (gw,) = torch.autograd.grad(yr, wr, dy.float().permute(0, 3, 1, 2))
    scale = gw.abs().max().item()
    assert (dw - gw).abs().max().item() <= 2e-3 * scale, ((dw - gw).abs().max().item(), scale)
    dw2 = torch.zeros_like(dw)
    ops.tc_conv_wgrad(x, Cin, dy, Cout, dw2, N, H, W, Cin, Cout, 3, workspace=ws)
    assert torch.equal(dw, dw2)
    monkeypatch.setenv("STARCOP_NO_WGRAD_HALO", "1")
    assert lib.sc_tc_wgrad_halo_supported(N, H, W, Cin, Cout) == 0
    dw3 = torch.zeros_like(dw)
    ops.tc_conv_wgrad(x, Cin, dy, Cout, dw3, N, H, W, Cin, Cout, 3)
    assert (dw - dw3).abs().max().item() <= 1e-4 * scale      # same bf16 products, fp32 accumulation order differs
    # channel-slice operands (explicit ld): x is a slice of a wider concat buffer
    monkeypatch.delenv("STARCOP_NO_WGRAD_HALO")
    wide = torch.randn(N, H, W, Cin + 16, device=DEV).to(torch.bfloat16)
    wide[..., 8:8 + Cin] = x
    dw4 = torch.zeros_like(dw)
    xs = wide[..., 8:]
    lib_ws = torch.empty(max(lib.sc_tc_conv_wgrad_workspace_bytes(N, H, W, Cin, Cout, 3, 3, 1), 4), dtype=torch.uint8, device=DEV)
    call("sc_tc_conv_wgrad", xs.data_ptr(), Cin + 16, dy.data_ptr(), Cout, dw4.data_ptr(), lib_ws.data_ptr(), N, H, W, Cin, Cout,
         3, 3, 1, st())
    assert torch.equal(dw, dw4)
