import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from starcop_b200 import synthetic
from starcop_b200.model_setup import get_model
from starcop_b200.settings import default_settings
dev = "cuda"
for size in (64, 128, 256, 512):
    b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synthetic.hyperstarcop_batch(2, size=size, seed=5).items()}
    torch.manual_seed(1234)
    mf = get_model(default_settings(pos_weight=1.0, compute_dtype="f32"), None).to(dev).train()
    with torch.no_grad():
        ref = mf(b["input"])
    for halo in ("1", ""):
        os.environ["STARCOP_NO_HALO"] = halo
        torch.manual_seed(1234)
        m = get_model(default_settings(pos_weight=1.0, compute_dtype="bf16"), None).to(dev).train()
        outs = []
        with torch.no_grad():
            for i in range(3):
                outs.append(m(b["input"]).clone())
        torch.cuda.synchronize()
        print(size, "no_halo" if halo else "halo", [f"{(o - ref).abs().max().item():.4f}" for o in outs],
              "finite", [bool(torch.isfinite(o).all()) for o in outs])
