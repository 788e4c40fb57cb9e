import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from starcop_b200 import mag1c, synthetic, _lib
t73 = synthetic.synthetic_template(73)
cube8, _, _ = synthetic.aviris_cube(2, size=512, bands=125, seed=1, template=t73)
c = torch.from_numpy(cube8).cuda()
lib = _lib.load()
for it in (0, 30):
    for _ in range(2): mag1c.mag1c_tiles(c, t73, slice(52, 125), num_iter=it)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 16)()
    lib.sc_debug_mag1c_clocks(buf)
    t = list(buf)
    names = ["setup+load", "xbar", "covariance", "inverse", "it0 to apply end", "iteration 1", "iterations 2..N"]
    print("num_iter", it, {n: t[i + 1] - t[i] for i, n in enumerate(names) if t[i + 1] > t[i]}, "total", t[7] - t[0])
