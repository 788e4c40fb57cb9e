import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time
from oracle import mag1c as omag
from starcop_b200 import mag1c, synthetic
t73 = np.load("tests/golden/ch4_template_aviris.npz")["template"][:, 1]
cube, _, alpha = synthetic.aviris_cube(1, size=128, bands=125, seed=9, template=t73)
sl = slice(52, 125)
mf, al = mag1c.mag1c_tiles(torch.from_numpy(cube).cuda(), t73, sl, num_iter=30)
m64, a64 = omag.mag1c_tile_columns(cube[0].astype(np.float64), t73, sl, num_iter=30)
m32, a32 = omag.mag1c_tile_columns(cube[0], t73, sl, num_iter=30)
g64, _ = mag1c.mag1c_tiles(torch.from_numpy(cube.astype(np.float64)).cuda(), t73, sl, num_iter=30)
sc = m64.abs().max().item()
print("scale", sc, "gpu32-vs-64", (mf[0].cpu().double()-m64).abs().max().item(), "cpu32-vs-64", (m32.double()-m64).abs().max().item(),
      "gpu64-vs-64", (g64[0].cpu()-m64).abs().max().item())
for it in (0, 1, 2, 5):
    a, _ = mag1c.mag1c_tiles(torch.from_numpy(cube).cuda(), t73, sl, num_iter=it)
    x = torch.as_tensor(np.ascontiguousarray(cube[0][:, :, sl])).permute(1, 0, 2).contiguous().double()
    b, _ = omag.acrwl1mf(x, torch.as_tensor(t73), num_iter=it)
    print(it, (a[0].cpu().double() - b[..., 0].T).abs().max().item())
# timing at the bench workload
cube8, _, _ = synthetic.aviris_cube(2, size=512, bands=125, seed=1, template=t73)
c = torch.from_numpy(cube8).cuda()
for _ in range(2): mag1c.mag1c_tiles(c, t73, sl)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): mag1c.mag1c_tiles(c, t73, sl)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
print(f"mag1c 2 tiles 512x512x73: {dt*1e3:.2f} ms -> {2/dt:.1f} tiles/s")
