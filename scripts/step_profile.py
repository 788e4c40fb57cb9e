"""Per-launch CUDA-event profile of one eager train step (starcop_b200/profiler.py), with shapes.
    python scripts/step_profile.py [--batch 16] [--size 512] [--filter bn_bwd]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from starcop_b200 import profiler, synthetic
from starcop_b200.model_setup import get_model
from starcop_b200.settings import default_settings
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--filter", default="")
ap.add_argument("--min-us", type=float, default=0.0)
a = ap.parse_args()
dev = "cuda"
torch.manual_seed(0)
m = get_model(default_settings(pos_weight=1.0, compute_dtype="bf16"), None).to(dev).train()
b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synthetic.hyperstarcop_batch(a.batch, size=a.size, seed=1).items()}
m.network._materialize()
m.network._engine.side_wgrad = False
for _ in range(2):
    m.train_step_fused(b)
SHAPE = {"sc_bn_bwd_reduce": lambda x: f"pooled={x[2]} N{x[12]} {x[13]}x{x[14]} C{x[15]} lddz{x[1]} ldy{x[4]}",
         "sc_bn_bwd_apply": lambda x: f"pooled={x[2]} N{x[17]} {x[18]}x{x[19]} C{x[20]}",
         "sc_bn_act": lambda x: f"up2={x[13]} res={int(bool(x[5]))} N{x[9]} {x[10]}x{x[11]} C{x[12]} ldz{x[8]}",
         "sc_bn_stats": lambda x: f"P{x[4]} C{x[5]}",
         "sc_tc_conv_fprop": lambda x: f"N{x[7]} {x[8]}x{x[9]} {x[10]}->{x[11]} k{x[12]} s{x[14]} acc{x[15]} stats={int(bool(x[5]))}",
         "sc_tc_conv_wgrad": lambda x: f"N{x[6]} {x[7]}x{x[8]} {x[9]}->{x[10]} k{x[11]} s{x[13]}",
         "sc_tc_conv3x3_halo": lambda x: f"N{x[7]} {x[8]}x{x[9]} {x[10]}->{x[11]} acc{x[12]}",
         "sc_dwconv_fprop": lambda x: f"N{x[10]} {x[11]}x{x[12]} C{x[13]} s{x[14]}",
         "sc_dwconv_dgrad": lambda x: f"N{x[5]} {x[6]}x{x[7]} C{x[8]} s{x[9]}",
         "sc_dwconv_wgrad": lambda x: f"N{x[9]} {x[10]}x{x[11]} C{x[12]} s{x[13]}"}
shapes = []
orig_enter = profiler.StepProfile.__enter__
with profiler.StepProfile() as prof:
    inner = profiler._lib.call
    def spy(name, *args):
        shapes.append(SHAPE[name](args) if name in SHAPE else "")
        inner(name, *args)
    profiler._lib.call = spy
    from starcop_b200 import engine
    engine.call = spy
    m.train_step_fused(b)
tot = sum(t for _, t, _, _ in prof.rows)
print(f"launches {len(prof.rows)}  sum {tot*1e3:.2f} ms")
for (n, t, fl, by), sh in zip(prof.rows, shapes):
    if a.filter in n and t * 1e6 >= a.min_us:
        print(f"{n:24s} {t*1e6:8.1f} us  {by/t/1e9 if t else 0:7.0f} GB/s  {fl/t/1e12 if t else 0:7.1f} TF/s  {sh}")
