"""Tensor-core resident mag1c kernel vs the fp64 streaming kernel (same fixed-point iteration, different algorithms):
errors, phase clocks of group 0, and timings at the bench shape.
    python scripts/diag_mag1c_tc.py"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from starcop_b200 import mag1c, synthetic, _lib
lib = _lib.load()
t73 = synthetic.synthetic_template(73)
sl = slice(52, 125)


def run(c, it, env=None):
    for k in ("STARCOP_MAG1C_STREAMING", "STARCOP_MAG1C_NO_TC"):
        os.environ.pop(k, None)
    if env:
        os.environ[env] = "1"
    out = mag1c.mag1c_tiles(c, t73, sl, num_iter=it)
    torch.cuda.synchronize()
    os.environ.pop(env or "x", None)
    return out


for size in (128, 512):
    cube, _, _ = synthetic.aviris_cube(2, size=size, bands=125, seed=11, template=t73)
    c = torch.from_numpy(cube).cuda()
    c64 = c.double()
    for it in (0, 1, 30):
        m_tc, a_tc = run(c, it)
        m_v2, a_v2 = run(c, it, "STARCOP_MAG1C_NO_TC")
        m_st, a_st = run(c, it, "STARCOP_MAG1C_STREAMING")
        m_64, a_64 = mag1c.mag1c_tiles(c64, t73, sl, num_iter=it)          # fp64 data, streaming kernel
        sc = m_64.abs().max().item()
        e = lambda a, b: (a.double() - b.double()).abs().max().item() / sc
        print(f"size {size} it {it}: scale {sc:.3g}  tc-vs-f64 {e(m_tc, m_64):.2e}  v2-vs-f64 {e(m_v2, m_64):.2e}  "
              f"stream32-vs-f64 {e(m_st, m_64):.2e}  albedo tc {((a_tc.double() - a_64) / a_64).abs().max().item():.2e} "
              f"finite {bool(torch.isfinite(m_tc).all())}")
cube, _, _ = synthetic.aviris_cube(2, size=512, bands=125, seed=1, template=t73)
c = torch.from_numpy(cube).cuda()
names = ["load", "centre", "mma", "assemble", "inverse", "rmf apply", "it0 (rest)", "iterations"]
for it in (0, 30):
    run(c, it)
    buf = (ctypes.c_longlong * 16)()
    lib.sc_debug_mag1c_clocks(buf)
    t = list(buf)
    print("num_iter", it, "clocks", [t[i] - t[0] for i in range(9)])
    if it:
        print("   iteration 1: start->vectors", t[9] - t[7], "matvec", t[10] - t[9], "scalars", t[11] - t[10], "apply", t[12] - t[11],
              "sums", t[13] - t[12], "v pass", t[14] - t[13])
for env in (None, "STARCOP_MAG1C_NO_TC"):
    for it in (0, 1, 5, 30):
        for _ in range(2):
            run(c, it, env)
        if env:
            os.environ[env] = "1"
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            mag1c.mag1c_tiles(c, t73, sl, num_iter=it)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        os.environ.pop(env or "x", None)
        print(f"{env or 'tc'} num_iter={it:2d}: {dt*1e3:.3f} ms for 2 tiles = {dt*5e5:.1f} us/tile")
