"""CUDA-event timing of the band-ratio product (exact 5/95 percentile select + apply) over T tiles of 512x512."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from starcop_b200 import features
for T in (8, 18, 64, 256):
    bg = torch.rand(T, 512, 512, device="cuda") + 0.5
    sig = bg * 0.8 + 0.01 * torch.randn_like(bg)
    for _ in range(3):
        features.ratio_2c_match_c_from_sums_outlier(bg, sig)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        features.ratio_2c_match_c_from_sums_outlier(bg, sig)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"T={T:4d}: {us:8.1f} us total  {us / T:7.2f} us/tile  {T * 512 * 512 * 12 / us / 1e3:7.0f} GB/s (12 B/px algorithmic)")
