import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from starcop_b200 import srf
dev = "cuda"; tiles, size, bands = 8, 512, 125
cube = torch.rand(tiles, size, size, bands, device=dev) + 0.5
centers = 380.0 + 17.0 * np.arange(bands); wl = np.arange(400.0, 2400.0, 2.0)
Wt = srf.srf_weight_table(wl, np.stack([np.exp(-0.5 * ((wl - (450 + 230 * k)) / 40.0) ** 2) for k in range(8)]), centers)
for _ in range(3): srf.transform_to_srf(cube, Wt, 0.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): srf.transform_to_srf(cube, Wt, 0.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{os.environ.get('STARCOP_SRF_KB','32')}KB ctas={os.environ.get('STARCOP_SRF_CTAS','2')} dbg={os.environ.get('STARCOP_SRF_DBG','0')}: {ms*1e3/tiles:.1f} us/tile  {tiles*size*size*(bands+8)*4/ms/1e6:.0f} GB/s")
