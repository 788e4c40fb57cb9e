"""Per-parameter gradient error of the CUDA engine (f32 / bf16) and of the fp32 CPU oracle, all
measured against the SAME network evaluated in float64 on the CPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.module import get_model as og
from starcop_b200 import synthetic
from starcop_b200.model_setup import get_model
from starcop_b200.settings import default_settings

size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
batch = synthetic.hyperstarcop_batch(B, size=size, seed=3)
dev = "cuda"

def grads_oracle(dt):
    torch.manual_seed(1234)
    o = og(default_settings(pos_weight=1.0)).train()
    if dt == torch.float64:
        o = o.double()
        b = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in batch.items()}
        # normalize_x casts to float: bypass by calling network directly on the float32-normalised input
        from oracle.normalizer import normalize_x
        from oracle import loss_metrics as lm
        x = normalize_x(batch["input"], o.input_products).double()
        logits = o.network(x)
        loss = torch.mean(lm.bce_with_logits_elementwise(logits, b["output"], o.pos_weight.double()) * b["weight_loss"])
    else:
        loss = o.training_step(batch, 0)
    loss.backward()
    return loss.item(), {n: p.grad.double() for n, p in o.network.named_parameters()}

def grads_gpu(cd):
    torch.manual_seed(1234)
    m = get_model(default_settings(pos_weight=1.0, compute_dtype=cd), None).to(dev).train()
    loss = m.training_step({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}, 0)
    loss.backward()
    return loss.item(), {n: p.grad.double().cpu() for n, p in m.network.named_parameters()}

l64, g64 = grads_oracle(torch.float64)
l32, g32 = grads_oracle(torch.float32)
lf, gf = grads_gpu("f32")
lb, gb = grads_gpu("bf16")
print(f"loss f64 {l64:.8f} cpu-f32 {l32:.8f} gpu-f32 {lf:.8f} gpu-bf16 {lb:.8f}")
def stats(g, r):
    rel = (g - r).abs().max().item() / max(r.abs().max().item(), 1e-30)
    cos = torch.nn.functional.cosine_similarity(g.flatten(), r.flatten(), dim=0).item()
    return rel, cos
print(f"{'param':55s} {'cpu32 rel':>10s} {'gpu32 rel':>10s} {'bf16 rel':>10s} {'bf16 cos':>9s}")
for n in g64:
    if n.endswith(".weight") and g64[n].dim() == 4 or "head" in n:
        a, _ = stats(g32[n], g64[n]); b, _ = stats(gf[n], g64[n]); c, cc = stats(gb[n], g64[n])
        print(f"{n:55s} {a:10.2e} {b:10.2e} {c:10.2e} {cc:9.4f}")
