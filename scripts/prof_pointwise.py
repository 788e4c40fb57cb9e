"""One 1x1 expansion conv (no statistics epilogue: the configuration the engine runs) for ncu / timing.
    python scripts/prof_pointwise.py [cin cout div]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from starcop_b200 import _lib
from starcop_b200._lib import call
lib = _lib.load()
cin, cout, div = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (16, 96, 2)))
B, S = 16, 512 // div
st = torch.cuda.current_stream().cuda_stream
xs = [torch.randn(B, S, S, cin, device="cuda").bfloat16() for _ in range(3)]
ys = [torch.empty(B, S, S, cout, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
w = torch.randn(cout, cin, 1, 1, device="cuda") * 0.05
cpad = lib.sc_tc_cin_pad(cin)
wb = torch.empty(cout * cpad, dtype=torch.bfloat16, device="cuda")
call("sc_tc_pack_weights", w.data_ptr(), wb.data_ptr(), cout, cin, 1, 1, 0, cpad, cout, st)
n = ctypes.c_int(0)
def run(i):
    call("sc_tc_conv_fprop", xs[i % 3].data_ptr(), cin, wb.data_ptr(), ys[i % 3].data_ptr(), cout, 0, ctypes.byref(n), B, S, S, cin, cout, 1, 1, 1, 0, st)
for i in range(3): run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10): run(i)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
print(f"{cin}->{cout} @{S}x{S}: {us:.1f} us  {B*S*S*(cin+cout)*2/us/1e3:.0f} GB/s")
