"""Runs only the dominant kernel (decoder.blocks.0.conv1 fprop, tcgen05) a few times: the target of
    ncu --set full --clock-control none --import-source on -k regex:tc_conv_fprop -s 2 -c 2 -o gpurun_out/prof_b0c1 \
        python scripts/profile_dominant.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from starcop_b200.model_setup import get_model
from starcop_b200.settings import default_settings
torch.manual_seed(0)
m = get_model(default_settings(compute_dtype="bf16"), None).cuda()
r = m.network.bench_dominant_kernel(16, 512, iters=4)
print(r)
