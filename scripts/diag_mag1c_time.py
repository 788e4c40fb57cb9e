import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from starcop_b200 import mag1c, synthetic
t73 = np.load("tests/golden/ch4_template_aviris.npz")["template"][:, 1]
cube8, _, _ = synthetic.aviris_cube(2, size=512, bands=125, seed=1, template=t73)
c = torch.from_numpy(cube8).cuda()
sl = slice(52, 125)
for it in (0, 1, 5, 30):
    for _ in range(2): mag1c.mag1c_tiles(c, t73, sl, num_iter=it)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): mag1c.mag1c_tiles(c, t73, sl, num_iter=it)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"num_iter={it:2d}: {dt*1e3:.2f} ms for 2 tiles")
