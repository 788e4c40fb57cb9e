import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from starcop_b200 import features
T = int(os.environ.get("T", "8"))
bg = torch.rand(T, 512, 512, device="cuda") + 0.5
sig = bg * 0.8 + 0.01 * torch.randn_like(bg)
for _ in range(2):
    features.ratio_2c_match_c_from_sums_outlier(bg, sig)
torch.cuda.synchronize()
