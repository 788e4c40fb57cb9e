"""fp32 parity diagnostics: per parameter, error of the CUDA fp32 engine and of the CPU fp32 oracle against the
float64 evaluation of the same network; prints the worst gpu/cpu ratios (the test allows 3x, floor 1e-4)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_unet import build_pair, f64_truth, rel_err, to_dev
from starcop_b200 import synthetic
for pw in (1.0, 15.0):
    oracle, model = build_pair(pw)
    batch = synthetic.hyperstarcop_batch(2, size=64, seed=3)
    oracle.train(); model.train()
    lo = oracle.training_step(batch, 0); lo.backward()
    lm_ = model.training_step(to_dev(batch), 0); lm_.backward()
    l64, g64 = f64_truth(pw, batch)
    rows = []
    for (n, po), (_, pm) in zip(oracle.network.named_parameters(), model.network.named_parameters()):
        e_gpu = rel_err(pm.grad.cpu().double(), g64[n]); e_cpu = rel_err(po.grad.double(), g64[n])
        rows.append((e_gpu / max(e_cpu, 1e-4 / 3), e_gpu, e_cpu, n))
    rows.sort(reverse=True)
    print(f"pw={pw}: loss gpu {lm_.item():.8f} cpu {lo.item():.8f} f64 {l64:.8f}; params {len(rows)}; ratio>3: {sum(r[0] > 3 for r in rows)}, >2: {sum(r[0] > 2 for r in rows)}, >1: {sum(r[0] > 1 for r in rows)}")
    for r in rows[:8]:
        print(f"   ratio {r[0]:6.2f}  gpu {r[1]:.3e}  cpu {r[2]:.3e}  {r[3]}")
