import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from starcop_b200 import mag1c, synthetic
t73 = synthetic.synthetic_template(73)
cube, _, _ = synthetic.aviris_cube(2, size=512, bands=125, seed=1, template=t73)
c = torch.from_numpy(cube).cuda()
for it in (int(os.environ.get("ITERS", "30")),):
    mag1c.mag1c_tiles(c, t73, slice(52, 125), num_iter=it)
torch.cuda.synchronize()
