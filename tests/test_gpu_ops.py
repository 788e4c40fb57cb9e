"""Per-kernel parity: each C-ABI entry point against plain PyTorch fp32 on the same seeded inputs."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from starcop_b200 import _lib, ops  # noqa: E402
from starcop_b200._lib import ACT_NONE, ACT_RELU, ACT_RELU6, SC_BF16, SC_F32, call, load  # noqa: E402

DEV = "cuda"
TDT = {SC_F32: torch.float32, SC_BF16: torch.bfloat16}


def st():
    return torch.cuda.current_stream().cuda_stream


def nhwc(x, dtype=SC_F32):
    return x.permute(0, 2, 3, 1).contiguous().to(TDT[dtype])


def nchw(x):
    return x.permute(0, 3, 1, 2).float()


def tol(dtype):
    return dict(rtol=2e-5, atol=2e-5) if dtype == SC_F32 else dict(rtol=3e-2, atol=3e-2)


@pytest.mark.parametrize("dtype", [SC_F32, SC_BF16])
@pytest.mark.parametrize("cin,cout,k,stride,hw", [(4, 32, 3, 2, 32), (16, 96, 1, 1, 16), (24, 144, 1, 1, 8),
                                                  (96, 24, 1, 1, 8), (40, 16, 3, 1, 16), (1376, 256, 3, 1, 4),
                                                  (32, 16, 3, 1, 32), (320, 1280, 1, 1, 2)])
def test_conv_fprop_wgrad_dgrad(dtype, cin, cout, k, stride, hw):
    torch.manual_seed(0)
    N = 2
    x = torch.randn(N, cin, hw, hw, device=DEV)
    w = torch.randn(cout, cin, k, k, device=DEV) / np.sqrt(cin * k * k)
    pad = k // 2
    xh = nhwc(x, dtype)
    if cin % 8 and dtype == SC_BF16:
        pytest.skip("bf16 activations use 8-channel padded inputs")
    xr, wr = xh.float().permute(0, 3, 1, 2).requires_grad_(True), w.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, stride=stride, padding=pad)
    ho = yr.shape[-1]
    wp = torch.empty(k * k * cin * cout, device=DEV)
    call("sc_pack_weights", w.data_ptr(), wp.data_ptr(), cout, cin, k, k, 0, st())
    y = torch.empty(N, ho, ho, cout, device=DEV, dtype=TDT[dtype])
    call("sc_conv_fprop", xh.data_ptr(), cin, wp.data_ptr(), 0, y.data_ptr(), cout, N, hw, hw, cin, cout, k, k,
         stride, pad, dtype, 0, st())
    assert torch.allclose(nchw(y), yr, **tol(dtype))
    # backward
    dy = torch.randn_like(yr)
    dyh = nhwc(dy, dtype)
    yr.backward(dyh.float().permute(0, 3, 1, 2))
    dw = torch.zeros_like(w)
    ops.conv_wgrad(xh, cin, dyh, cout, dw, N, hw, hw, cin, cout, k, stride, pad, dtype)
    dw2 = torch.zeros_like(w)
    ops.conv_wgrad(xh, cin, dyh, cout, dw2, N, hw, hw, cin, cout, k, stride, pad, dtype)
    assert torch.equal(dw, dw2)                      # deterministic pixel splits: bit-identical from run to run
    scale = wr.grad.abs().max().item()
    assert (dw - wr.grad).abs().max().item() <= (1e-4 if dtype == SC_F32 else 2e-2) * scale
    if stride == 1 and cout % 8 == 0:
        wpt = torch.empty_like(wp)
        call("sc_pack_weights", w.data_ptr(), wpt.data_ptr(), cout, cin, k, k, 1, st())
        dx = torch.zeros(N, hw, hw, cin, device=DEV, dtype=TDT[dtype])
        call("sc_conv_fprop", dyh.data_ptr(), cout, wpt.data_ptr(), 0, dx.data_ptr(), cin, N, ho, ho, cout, cin,
             k, k, 1, pad, dtype, 0, st())
        s = xr.grad.abs().max().item()
        assert (nchw(dx) - xr.grad).abs().max().item() <= (1e-4 if dtype == SC_F32 else 3e-2) * s
        # accumulate flag
        call("sc_conv_fprop", dyh.data_ptr(), cout, wpt.data_ptr(), 0, dx.data_ptr(), cin, N, ho, ho, cout, cin,
             k, k, 1, pad, dtype, 1, st())
        assert (nchw(dx) - 2 * xr.grad).abs().max().item() <= (2e-4 if dtype == SC_F32 else 6e-2) * s


@pytest.mark.parametrize("dtype", [SC_F32, SC_BF16])
@pytest.mark.parametrize("c,stride,hw,fused", [(32, 1, 16, False), (96, 2, 16, True), (144, 1, 8, True), (960, 1, 4, True),
                                               (48, 1, 37, True), (24, 2, 31, False), (64, 2, 64, True), (192, 1, 40, True)])
def test_depthwise(dtype, c, stride, hw, fused):
    torch.manual_seed(1)
    N = 2
    x = torch.randn(N, c, hw, hw, device=DEV)
    w = torch.randn(c, 1, 3, 3, device=DEV) / 3
    scale, shift = torch.rand(c, device=DEV) + 0.5, torch.randn(c, device=DEV)
    xh = nhwc(x, dtype)
    xr = xh.float().permute(0, 3, 1, 2).requires_grad_(True)
    xin = F.relu6(xr * scale[None, :, None, None] + shift[None, :, None, None]) if fused else xr
    wr = w.clone().requires_grad_(True)
    yr = F.conv2d(xin, wr, stride=stride, padding=1, groups=c)
    ho = yr.shape[-1]
    y = torch.empty(N, ho, ho, c, device=DEV, dtype=TDT[dtype])
    sp, hp, act = (scale.data_ptr(), shift.data_ptr(), ACT_RELU6) if fused else (0, 0, ACT_NONE)
    part = torch.full((load().sc_bn_partials_bytes(c) // 8,), float("nan"), dtype=torch.float64, device=DEV)
    nrows = ctypes.c_int(0)
    call("sc_dwconv_fprop", xh.data_ptr(), c, sp, hp, act, w.data_ptr(), y.data_ptr(), c, part.data_ptr(),
         ctypes.byref(nrows), N, hw, hw, c, stride, dtype, st())
    assert torch.allclose(nchw(y), yr, **tol(dtype))
    # fused BatchNorm partial rows = sum / sum of squares of the STORED outputs
    rows = part[:nrows.value * 2 * c].view(nrows.value, 2 * c).sum(0)
    yf = y.float().reshape(-1, c).double()
    assert torch.allclose(rows[:c], yf.sum(0), rtol=1e-5, atol=1e-4)
    assert torch.allclose(rows[c:], (yf * yf).sum(0), rtol=1e-5, atol=1e-4)
    y2 = torch.empty_like(y)
    call("sc_dwconv_fprop", xh.data_ptr(), c, sp, hp, act, w.data_ptr(), y2.data_ptr(), c, 0, 0, N, hw, hw, c, stride, dtype, st())
    assert torch.equal(y, y2)
    dy = nhwc(torch.randn_like(yr), dtype)
    gin = torch.autograd.grad(yr, [xin, wr], dy.float().permute(0, 3, 1, 2))
    dx = torch.empty(N, hw, hw, c, device=DEV, dtype=TDT[dtype])
    call("sc_dwconv_dgrad", dy.data_ptr(), c, w.data_ptr(), dx.data_ptr(), c, N, hw, hw, c, stride, dtype, st())
    assert torch.allclose(nchw(dx), gin[0], **tol(dtype))
    dw = torch.zeros_like(w)
    ws = torch.empty(load().sc_dwconv_wgrad_workspace_bytes(c) // 4, device=DEV)
    call("sc_dwconv_wgrad", xh.data_ptr(), c, sp, hp, act, dy.data_ptr(), c, dw.data_ptr(), ws.data_ptr(), N, hw, hw, c, stride, dtype, st())
    assert (dw - gin[1]).abs().max().item() <= (1e-4 if dtype == SC_F32 else 2e-2) * gin[1].abs().max().item()


@pytest.mark.parametrize("dtype", [SC_F32, SC_BF16])
@pytest.mark.parametrize("c,hw,act,up2,res", [(16, 16, ACT_RELU6, False, False), (96, 8, ACT_NONE, False, True),
                                              (1280, 2, ACT_RELU6, True, False), (32, 16, ACT_RELU, True, False),
                                              (2064, 2, ACT_RELU, False, False),
                                              # >= 2^18 pixels, contiguous, C <= 256: the bulk-copy ("flat")
                                              # backward-reduce kernel, with a ragged last chunk
                                              (16, 304, ACT_RELU, False, False), (64, 300, ACT_RELU6, False, False),
                                              # the same kernel when C / 8 does not divide its block (idle threads):
                                              # the encoder's expanded tensors
                                              (96, 296, ACT_RELU6, False, False), (144, 300, ACT_RELU6, False, False)])
def test_batchnorm_train_forward_backward(dtype, c, hw, act, up2, res):
    torch.manual_seed(2)
    N = 3
    y = torch.randn(N, c, hw, hw, device=DEV) * 2 + 0.5
    yh = nhwc(y, dtype)
    yr = yh.float().permute(0, 3, 1, 2).requires_grad_(True)
    bn = torch.nn.BatchNorm2d(c).to(DEV).train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_()
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    r = torch.randn(N, c, hw, hw, device=DEV)
    rh = nhwc(r, dtype)
    zr = bn(yr)
    zr = {ACT_NONE: lambda t: t, ACT_RELU: F.relu, ACT_RELU6: F.relu6}[act](zr)
    if res:
        zr = zr + rh.float().permute(0, 3, 1, 2)
    if up2:
        zr = F.interpolate(zr, scale_factor=2, mode="nearest")
    P = N * hw * hw
    import ctypes
    sums = torch.full((load().sc_bn_partials_bytes(c) // 8,), float("nan"), dtype=torch.float64, device=DEV)
    nrows = ctypes.c_int(0)
    call("sc_bn_stats", yh.data_ptr(), c, sums.data_ptr(), ctypes.byref(nrows), P, c, dtype, st())
    assert 1 <= nrows.value <= 296
    scale, shift, mean, invstd = (torch.empty(c, device=DEV) for _ in range(4))
    call("sc_bn_finalize", sums.data_ptr(), nrows.value, P, c, bn.weight.data_ptr(), bn.bias.data_ptr(), rm.data_ptr(), rv.data_ptr(),
         0.1, 1e-5, 1, scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), st())
    assert torch.allclose(rm, bn.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(rv, bn.running_var, rtol=1e-5, atol=1e-6)
    f = 2 if up2 else 1
    z = torch.empty(N, hw * f, hw * f, c, device=DEV, dtype=TDT[dtype])
    call("sc_bn_act", yh.data_ptr(), c, scale.data_ptr(), shift.data_ptr(), act, rh.data_ptr() if res else 0, c,
         z.data_ptr(), c, N, hw, hw, c, int(up2), dtype, st())
    assert torch.allclose(nchw(z), zr, **tol(dtype))
    # backward
    dz = nhwc(torch.randn_like(zr), dtype)
    gy, gw, gb = torch.autograd.grad(zr, [yr, bn.weight, bn.bias], dz.float().permute(0, 3, 1, 2))
    red = torch.full((load().sc_bn_partials_bytes(c) // 8,), float("nan"), dtype=torch.float64, device=DEV)
    call("sc_bn_bwd_reduce", dz.data_ptr(), c, int(up2), yh.data_ptr(), c, scale.data_ptr(), shift.data_ptr(),
         mean.data_ptr(), invstd.data_ptr(), act, red.data_ptr(), ctypes.byref(nrows), N, hw, hw, c, dtype, st())
    dy = torch.empty(N, hw, hw, c, device=DEV, dtype=TDT[dtype])
    dg, db = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    call("sc_bn_bwd_apply", dz.data_ptr(), c, int(up2), yh.data_ptr(), c, scale.data_ptr(), shift.data_ptr(),
         mean.data_ptr(), invstd.data_ptr(), bn.weight.data_ptr(), act, red.data_ptr(), nrows.value, dy.data_ptr(), c,
         dg.data_ptr(), db.data_ptr(), N, hw, hw, c, dtype, st())
    t = dict(rtol=1e-4, atol=1e-4) if dtype == SC_F32 else dict(rtol=5e-2, atol=5e-2)
    assert torch.allclose(nchw(dy), gy, **t)
    assert torch.allclose(dg, gw, rtol=1e-3, atol=1e-3 * gw.abs().max().item())
    assert torch.allclose(db, gb, rtol=1e-3, atol=1e-3 * gb.abs().max().item())


def test_bn_eval_uses_running_stats():
    c = 24
    rm, rv = torch.randn(c, device=DEV), torch.rand(c, device=DEV) + 0.5
    g, b = torch.rand(c, device=DEV) + 0.5, torch.randn(c, device=DEV)
    scale, shift = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
    call("sc_bn_finalize", 0, 0, 100, c, g.data_ptr(), b.data_ptr(), rm.data_ptr(), rv.data_ptr(), 0.1, 1e-5, 0,
         scale.data_ptr(), shift.data_ptr(), 0, 0, st())
    x = torch.randn(2, c, 4, 4, device=DEV)
    ref = F.batch_norm(x, rm, rv, g, b, False, 0.1, 1e-5)
    assert torch.allclose(x * scale[None, :, None, None] + shift[None, :, None, None], ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("dtype", [SC_F32, SC_BF16])
def test_head(dtype):
    torch.manual_seed(3)
    N, C, hw = 2, 16, 32
    x = nhwc(torch.randn(N, C, hw, hw, device=DEV), dtype)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    conv = torch.nn.Conv2d(C, 1, 3, padding=1).to(DEV)
    yr = conv(xr)
    lg = torch.empty(N, 1, hw, hw, device=DEV)
    call("sc_head_fprop", x.data_ptr(), C, conv.weight.data_ptr(), conv.bias.data_ptr(), lg.data_ptr(), N, hw, hw, C, dtype, st())
    assert torch.allclose(lg, yr, rtol=1e-5, atol=1e-5)
    dl = torch.randn_like(yr)
    gx, gw, gb = torch.autograd.grad(yr, [xr, conv.weight, conv.bias], dl)
    dx = torch.empty(N, hw, hw, C, device=DEV, dtype=TDT[dtype])
    call("sc_head_bwd", x.data_ptr(), C, conv.weight.data_ptr(), dl.data_ptr(), dx.data_ptr(), C, N, hw, hw, C, dtype, st())
    assert torch.allclose(nchw(dx), gx, **tol(dtype))


@pytest.mark.parametrize("h,w", [(8, 32), (40, 24), (19, 37), (64, 96), (7, 5)])
def test_head_tiled_kernels_ragged_sizes(h, w):
    """the shared-memory tiled head kernels (16 bf16 channels): tiles of 8 x 32 outputs with ragged borders, several
    tiles per CTA (the cp.async ring wraps), against torch's Conv2d on the same bf16 inputs"""
    torch.manual_seed(11)
    N, C = 3, 16
    x = nhwc(torch.randn(N, C, h, w, device=DEV), SC_BF16)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    conv = torch.nn.Conv2d(C, 1, 3, padding=1).to(DEV)
    yr = conv(xr)
    lg = torch.full((N, 1, h, w), float("nan"), device=DEV)
    call("sc_head_fprop", x.data_ptr(), C, conv.weight.data_ptr(), conv.bias.data_ptr(), lg.data_ptr(), N, h, w, C, SC_BF16, st())
    assert torch.allclose(lg, yr, rtol=1e-5, atol=1e-5)
    dl = torch.randn_like(yr)
    gx, = torch.autograd.grad(yr, [xr], dl)
    dx = torch.full((N, h, w, C), float("nan"), device=DEV, dtype=TDT[SC_BF16])
    call("sc_head_bwd", x.data_ptr(), C, conv.weight.data_ptr(), dl.data_ptr(), dx.data_ptr(), C, N, h, w, C, SC_BF16, st())
    assert torch.allclose(nchw(dx), gx, **tol(SC_BF16))


@pytest.mark.parametrize("dtype", [SC_F32, SC_BF16])
@pytest.mark.parametrize("C,h,w", [(16, 32, 32), (16, 40, 24), (8, 19, 37), (64, 16, 16), (32, 64, 48)])
def test_head_wgrad_tiled(dtype, C, h, w):
    """weight / bias gradient of the head through the TMA tile kernel (ragged tiles, several channel-block widths);
    dw / dbias are accumulated onto what is there"""
    torch.manual_seed(4)
    N = 3
    x = nhwc(torch.randn(N, C, h, w, device=DEV), dtype)
    xr = x.float().permute(0, 3, 1, 2)
    conv = torch.nn.Conv2d(C, 1, 3, padding=1).to(DEV)
    yr = conv(xr)
    dl = torch.randn_like(yr)
    gw, gb = torch.autograd.grad(yr, [conv.weight, conv.bias], dl)
    dw, db = torch.ones_like(conv.weight), torch.full_like(conv.bias, 2.0)
    ws = torch.empty(load().sc_head_wgrad_workspace_bytes(C) // 4, device=DEV)
    call("sc_head_wgrad_tiled", x.data_ptr(), C, dl.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), N, h, w, C, dtype, st())
    scale = gw.abs().max().item()
    assert (dw - 1.0 - gw).abs().max().item() <= 2e-5 * scale + 1e-4
    assert abs(db.item() - 2.0 - gb.item()) <= 1e-4 * max(1.0, abs(gb.item()))


def test_add_into_pooled_and_accumulate():
    N, C, hw = 2, 24, 6
    a = torch.randn(N, 2 * hw, 2 * hw, C + 8, device=DEV)
    out = torch.randn(N, hw, hw, C, device=DEV)
    ref = out + F.avg_pool2d(a[..., 8:].permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1) * 4
    call("sc_add_into", a.data_ptr() + 8 * 4, C + 8, 1, out.data_ptr(), C, 1, N, hw, hw, C, SC_F32, st())
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5)


def test_adam_matches_torch():
    torch.manual_seed(4)
    n = 10007
    p = torch.randn(n, device=DEV)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-4)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 6):
        g = torch.randn(n, device=DEV)
        pr.grad = g.clone()
        opt.step()
        call("sc_adam_step", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-4, 0.9, 0.999, 1e-8, step, 1.0, st())
    assert torch.allclose(p, pr.detach(), rtol=1e-6, atol=1e-7)


def test_normalize_pack_bit_exact(golden):
    from starcop_b200.normalizer import DataNormalizer
    from starcop_b200.settings import default_settings
    import warnings
    g = golden("normalizer.npz")
    prods = {"hyper": ["mag1c", "TOA_AVIRIS_640nm", "TOA_AVIRIS_550nm", "TOA_AVIRIS_460nm"],
             "multi": ["ratio_wv3_B7_B5_varon21_sum_c_out", "TOA_WV3_SWIR1",
                       "ratio_wv3_B8_B8MLR_SanchezGarcia22_simplediv", "unknown_product"]}
    for tag, p in prods.items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            dn = DataNormalizer(default_settings(input_products=p)).to(DEV)
        y = dn.normalize_x(torch.from_numpy(g[f"{tag}_x"]).to(DEV))
        assert np.array_equal(y.cpu().numpy(), g[f"{tag}_y"]), tag      # reference output, bit exact


def test_bce_fused_edge_cases(golden):
    g = golden("model_module.npz")
    x, y, w = (torch.from_numpy(g[k]).to(DEV) for k in ("edge_logits", "edge_y", "edge_w"))
    n = x.numel()
    for pw in (1, 15):
        loss = ops.bce_loss_buffer(1, n, DEV)
        grad, lpx, pred = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        pb, diff = torch.empty(n, dtype=torch.long, device=DEV), torch.empty(n, dtype=torch.long, device=DEV)
        cm, cms = torch.zeros(4, dtype=torch.long, device=DEV), torch.zeros(4, dtype=torch.long, device=DEV)
        call("sc_bce_fused", x.data_ptr(), y.data_ptr(), w.data_ptr(), float(pw), 1, n, 1.0 / n, loss.data_ptr(),
             grad.data_ptr(), cm.data_ptr(), 0, cms.data_ptr(), 0, pred.data_ptr(), lpx.data_ptr(), 0, pb.data_ptr(),
             diff.data_ptr(), st())
        assert np.allclose(lpx.cpu().numpy(), g[f"edge_pw{pw}_loss"], rtol=2e-6, atol=1e-30)
        assert np.allclose(grad.cpu().numpy(), g[f"edge_pw{pw}_grad"], rtol=1e-5, atol=1e-12)
        assert np.array_equal(pb.cpu().numpy(), g["edge_pred_sigmoid"])           # sigmoid(x) > .5, bit exact
        pv = g["edge_pred_val"]; yy = g["edge_y"].astype(np.int64)
        ref_cm = np.bincount(2 * yy + pv, minlength=4)
        assert np.array_equal(cm.cpu().numpy(), ref_cm)                             # logits >= 0, bit exact
        assert np.array_equal(cms.cpu().numpy(), np.bincount(2 * yy + g["edge_pred_sigmoid"], minlength=4))
        assert np.array_equal(diff.cpu().numpy(), 2 * g["edge_pred_sigmoid"] + (g["edge_y"] == 1))
        ref_loss = (g[f"edge_pw{pw}_loss"].astype(np.float64) * g["edge_w"]).sum()
        assert abs(loss[0].item() - ref_loss) <= 1e-6 * abs(ref_loss)
        assert loss[1].item() == 0.0                     # the ticket word is back at zero


def test_threshold_opening_matches_oracle():
    from oracle import morphology
    torch.manual_seed(5)
    pred = torch.rand(3, 1, 48, 40, device=DEV) * 1000
    out = torch.empty(3, 1, 48, 40, dtype=torch.long, device=DEV)
    scratch = torch.empty(3 * 48 * 40, dtype=torch.uint8, device=DEV)
    call("sc_threshold_opening", pred.data_ptr(), 500.0, out.data_ptr(), scratch.data_ptr(), 3, 48, 40, st())
    ref = morphology.apply_threshold(pred.cpu(), 500.0)
    assert torch.equal(out.cpu(), ref)


def test_weight_mag1c(golden):
    g = golden("features.npz")
    m = torch.from_numpy(g["mag1c"]).to(DEV)
    o = torch.empty_like(m)
    call("sc_weight_mag1c", m.data_ptr(), o.data_ptr(), m.numel(), st())
    assert np.array_equal(o.cpu().numpy(), g["weight_mag1c"])
