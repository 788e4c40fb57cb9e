"""Per-layer timing of the tcgen05 conv kernels at the bench workload (bs tiles of size x size):
every dense conv of the HyperSTARCOP U-Net, fprop / dgrad / wgrad, CUDA-event timed, with the
achieved TFLOP/s and the bf16 activation GB/s each launch moves (min traffic: in + out once).

    python scripts/layer_bench.py [--batch 16] [--size 512] [--only b0c1] [--iters 10]
"""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from starcop_b200 import _lib, ops
from starcop_b200._lib import call

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--only", default="")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--ops", default="fprop,dgrad,wgrad")
a = ap.parse_args()
lib = _lib.load()
dev = "cuda"
S, B = a.size, a.batch
# name, Cin, Cout, k, spatial divisor
layers = [("stem", 4, 32, 3, 1), ("f1.pw", 32, 16, 1, 2), ("f2.exp", 16, 96, 1, 2), ("f2.pw", 96, 24, 1, 4),
          ("f3.exp", 24, 144, 1, 4), ("f4.pw", 144, 32, 1, 8), ("f5.exp", 32, 192, 1, 8), ("f7.pw", 192, 64, 1, 16),
          ("f8.exp", 64, 384, 1, 16), ("f11.pw", 384, 96, 1, 16), ("f12.exp", 96, 576, 1, 16), ("f14.pw", 576, 160, 1, 32),
          ("f15.exp", 160, 960, 1, 32), ("f17.pw", 960, 320, 1, 32), ("f18", 320, 1280, 1, 32),
          ("b0c1", 1376, 256, 3, 16), ("b0c2", 256, 256, 3, 16), ("b1c1", 288, 128, 3, 8), ("b1c2", 128, 128, 3, 8),
          ("b2c1", 152, 64, 3, 4), ("b2c2", 64, 64, 3, 4), ("b3c1", 80, 32, 3, 2), ("b3c2", 32, 32, 3, 2),
          ("b4c1", 32, 16, 3, 1), ("b4c2", 16, 16, 3, 1)]
st = torch.cuda.current_stream().cuda_stream
NB = 3
def timeit(fn):
    for i in range(2): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters * 1e3   # us
print(f"{'layer':8s} {'op':6s} {'us':>9s} {'TFLOP/s':>9s} {'GB/s':>8s}  shape")
tot = {}
for name, cin, cout, k, div in layers:
    if a.only and a.only not in name: continue
    stride = 2 if name == "stem" else 1
    H = W = S // div
    Ho, Wo = H // stride, W // stride
    ldx = 8 if cin < 8 else cin
    xs = [torch.randn(B, H, W, ldx, device=dev).to(torch.bfloat16) for _ in range(NB)]
    ys = [torch.randn(B, Ho, Wo, cout, device=dev).to(torch.bfloat16) for _ in range(NB)]
    w = torch.randn(cout, cin, k, k, device=dev) * 0.05
    cpad = lib.sc_tc_cin_pad(cin)
    wb = torch.empty(cout * k * k * cpad, dtype=torch.bfloat16, device=dev)
    call("sc_tc_pack_weights", w.data_ptr(), wb.data_ptr(), cout, cin, k, k, 0, cpad, cout, st)
    cpad2 = lib.sc_tc_cin_pad(cout)
    wt = torch.empty(cin * k * k * cpad2, dtype=torch.bfloat16, device=dev)
    call("sc_tc_pack_weights", w.data_ptr(), wt.data_ptr(), cout, cin, k, k, 1, cin, cpad2, st)
    part = torch.empty(lib.sc_bn_partials_bytes(cout) // 8, dtype=torch.float64, device=dev)
    dw = torch.zeros_like(w)
    n = ctypes.c_int(0)
    flops = 2.0 * B * Ho * Wo * k * k * cin * cout
    byts = 2.0 * B * (H * W * cin + Ho * Wo * cout)
    ws = [None]
    ops_ = {
        "fprop": lambda i: call("sc_tc_conv_fprop", xs[i % NB].data_ptr(), ldx, wb.data_ptr(), ys[i % NB].data_ptr(), cout,
                                part.data_ptr(), ctypes.byref(n), B, H, W, cin, cout, k, k, stride, 0, st),
        "dgrad": lambda i: call("sc_tc_conv_fprop", ys[i % NB].data_ptr(), cout, wt.data_ptr(), xs[i % NB].data_ptr(), ldx,
                                0, 0, B, Ho, Wo, cout, cin, k, k, 1, 0, st),
        "wgrad": lambda i: ws.__setitem__(0, ops.tc_conv_wgrad(xs[i % NB], ldx, ys[i % NB], cout, dw, B, H, W, cin, cout, k,
                                                               stride, workspace=ws[0])),
    }
    halo = k == 3 and stride == 1 and lib.sc_tc_halo_supported(cin, cout) and lib.sc_tc_halo_supported(cout, cin)
    if halo:
        hp, hp2 = lib.sc_tc_halo_cin_pad(cin), lib.sc_tc_halo_cin_pad(cout)
        wbh = torch.empty(cout * 9 * hp, dtype=torch.bfloat16, device=dev)
        call("sc_tc_pack_weights", w.data_ptr(), wbh.data_ptr(), cout, cin, 3, 3, 0, hp, cout, st)
        wth = torch.empty(cin * 9 * hp2, dtype=torch.bfloat16, device=dev)
        call("sc_tc_pack_weights", w.data_ptr(), wth.data_ptr(), cout, cin, 3, 3, 1, cin, hp2, st)
        ops_["fprop_halo"] = lambda i: call("sc_tc_conv3x3_halo", xs[i % NB].data_ptr(), ldx, wbh.data_ptr(), ys[i % NB].data_ptr(),
                                           cout, part.data_ptr(), ctypes.byref(n), B, H, W, cin, cout, 0, st)
        ops_["dgrad_halo"] = lambda i: call("sc_tc_conv3x3_halo", ys[i % NB].data_ptr(), cout, wth.data_ptr(), xs[i % NB].data_ptr(),
                                           ldx, 0, 0, B, H, W, cout, cin, 0, st)
    for op in a.ops.split(",") + (["fprop_halo", "dgrad_halo"] if halo else []):
        if op == "dgrad" and (stride != 1 or cin % 8): continue
        us = timeit(ops_[op])
        tot[op] = tot.get(op, 0) + us
        print(f"{name:8s} {op:6s} {us:9.1f} {flops / us / 1e6:9.1f} {byts / us / 1e3:8.0f}  {cin}->{cout} k{k} @{H}x{W}")
print("totals (us):", {k: round(v) for k, v in tot.items()})
