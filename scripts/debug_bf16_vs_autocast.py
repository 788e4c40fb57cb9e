"""How far is bf16 storage from fp32 on this network?  ours(bf16) and PyTorch autocast(bf16) vs PyTorch fp32,
train-mode (batch statistics) and eval-mode logits, 2 tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.module import get_model as oracle_get_model
from oracle.normalizer import normalize_x
from starcop_b200 import synthetic
from starcop_b200.model_setup import get_model
from starcop_b200.settings import default_settings
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
for size in (128, 512):
    batch = synthetic.hyperstarcop_batch(2, size=size, seed=5)
    b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    torch.manual_seed(1234)
    o = oracle_get_model(default_settings(pos_weight=1.0)).to(dev).train()
    x = normalize_x(batch["input"], o.input_products).to(dev)
    torch.manual_seed(1234)
    m = get_model(default_settings(pos_weight=1.0, compute_dtype="bf16"), None).to(dev).train()
    for steps in (0, 3):
        if steps:
            opt = torch.optim.Adam(o.network.parameters(), 1e-3)
            for i in range(steps):
                bb = synthetic.hyperstarcop_batch(2, size=size, seed=50 + i)
                from oracle import loss_metrics as lm
                opt.zero_grad()
                lg = o.network(normalize_x(bb["input"], o.input_products).to(dev))
                l = torch.mean(lm.bce_with_logits_elementwise(lg, bb["output"].to(dev), o.pos_weight.to(dev)) * bb["weight_loss"].to(dev))
                l.backward(); opt.step()
            sd = {k: v for k, v in o.state_dict().items()}
            m.load_state_dict({k: v for k, v in sd.items() if k in m.state_dict()}, strict=False)
        for mode in ("train", "eval"):
            o.train(mode == "train"); m.train(mode == "train")
            with torch.no_grad():
                # running stats are touched by train-mode forwards: snapshot / restore so all three see the same state
                snap = {k: v.clone() for k, v in o.network.state_dict().items()}
                ref = o.network(x)
                o.network.load_state_dict(snap)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    ac = o.network(x).float()
                o.network.load_state_dict(snap)
                mine = m(b["input"])
                m.network.load_state_dict(snap, strict=False)
            f = lambda a: f"max {(a - ref).abs().max().item():.4f} mean {(a - ref).abs().mean().item():.5f} sig_max {(torch.sigmoid(a) - torch.sigmoid(ref)).abs().max().item():.4f} flips {int(((a >= 0) != (ref >= 0)).sum())}"
            print(f"size {size} after {steps} steps {mode:5s} ref std {ref.std().item():.4f} | autocast: {f(ac)} | ours: {f(mine)}")
