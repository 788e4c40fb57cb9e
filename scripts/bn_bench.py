"""CUDA-event timing of the BatchNorm backward reduce at the decoder's full-resolution shapes (bs 16)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from starcop_b200 import _lib
from starcop_b200._lib import call, ACT_RELU, SC_BF16
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
for C, hw in ((16, 512), (32, 256), (64, 128)):
    N = 16
    ys = [torch.randn(N, hw, hw, C, device="cuda").bfloat16() for _ in range(3)]
    dzs = [torch.randn(N, hw, hw, C, device="cuda").bfloat16() for _ in range(3)]
    v = [torch.rand(C, device="cuda") + 0.5 for _ in range(4)]
    red = torch.empty(lib.sc_bn_partials_bytes(C) // 8, dtype=torch.float64, device="cuda")
    n = ctypes.c_int(0)
    def run(i):
        call("sc_bn_bwd_reduce", dzs[i % 3].data_ptr(), C, 0, ys[i % 3].data_ptr(), C, v[0].data_ptr(), v[1].data_ptr(),
             v[2].data_ptr(), v[3].data_ptr(), ACT_RELU, red.data_ptr(), ctypes.byref(n), N, hw, hw, C, SC_BF16, st)
    for i in range(3): run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(12): run(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 12 * 1e3
    print(f"C={C:3d} {hw}x{hw}: {us:7.1f} us  {2 * N * hw * hw * C * 2 / us / 1e3:6.0f} GB/s  rows={n.value}")
