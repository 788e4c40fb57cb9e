"""CUDA-event timing of the depthwise 3x3 kernels at the bench workload (bs tiles of size x size): every
depthwise layer of the MobileNetV2 encoder, fprop (with fused BN+ReLU6 on load and BN statistics of the
outputs) / dgrad / wgrad, with the bf16 GB/s each launch moves (min traffic: input + output once).

    python scripts/dw_bench.py [--batch 16] [--size 512]
"""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from starcop_b200 import _lib
from starcop_b200._lib import call, ACT_RELU6, SC_BF16

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--only", default="")
a = ap.parse_args()
lib = _lib.load()
B, S = a.batch, a.size
# name, C, input spatial divisor, stride
layers = [("f1.dw", 32, 2, 1), ("f2.dw", 96, 2, 2), ("f3.dw", 144, 4, 1), ("f4.dw", 144, 4, 2), ("f5.dw", 192, 8, 1),
          ("f7.dw", 192, 8, 2), ("f8.dw", 384, 16, 1), ("f12.dw", 576, 16, 1), ("f14.dw", 576, 16, 2), ("f15.dw", 960, 32, 1)]
st = torch.cuda.current_stream().cuda_stream
NB = 3


def timeit(fn):
    for i in range(2): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters * 1e3


tot = {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0}
print(f"{'layer':8s} {'op':6s} {'us':>8s} {'GB/s':>8s}  shape")
for name, C, div, s in layers:
    if a.only and name not in a.only.split(","):
        continue
    H = S // div
    Ho = H // s
    xs = [torch.randn(B, H, H, C, device="cuda").bfloat16() for _ in range(NB)]      # rotating buffers: cold L2
    ys = [torch.empty(B, Ho, Ho, C, device="cuda", dtype=torch.bfloat16) for _ in range(NB)]
    dxs = [torch.empty(B, H, H, C, device="cuda", dtype=torch.bfloat16) for _ in range(NB)]
    w = torch.randn(C, 1, 3, 3, device="cuda") / 3
    dw = torch.zeros_like(w)
    sc, sh = torch.rand(C, device="cuda") + .5, torch.randn(C, device="cuda")
    part = torch.empty(lib.sc_bn_partials_bytes(C) // 8, dtype=torch.float64, device="cuda")
    ws = torch.empty(lib.sc_dwconv_wgrad_workspace_bytes(C) // 4, device="cuda")
    n = ctypes.c_int(0)
    byts = (B * H * H * C + B * Ho * Ho * C) * 2
    ops = {
        "fprop": lambda i: call("sc_dwconv_fprop", xs[i % NB].data_ptr(), C, sc.data_ptr(), sh.data_ptr(), ACT_RELU6, w.data_ptr(),
                                ys[i % NB].data_ptr(), C, part.data_ptr(), ctypes.byref(n), B, H, H, C, s, SC_BF16, st),
        "dgrad": lambda i: call("sc_dwconv_dgrad", ys[i % NB].data_ptr(), C, w.data_ptr(), dxs[i % NB].data_ptr(), C, B, H, H, C, s, SC_BF16, st),
        "wgrad": lambda i: call("sc_dwconv_wgrad", xs[i % NB].data_ptr(), C, sc.data_ptr(), sh.data_ptr(), ACT_RELU6, ys[i % NB].data_ptr(), C,
                                dw.data_ptr(), ws.data_ptr(), B, H, H, C, s, SC_BF16, st),
    }
    for op, fn in ops.items():
        us = timeit(fn)
        tot[op] += us
        print(f"{name:8s} {op:6s} {us:8.1f} {byts / us / 1e3:8.0f}  C={C} {H}x{H} s{s}")
print("totals (us, one launch per distinct shape):", {k: round(v) for k, v in tot.items()})
