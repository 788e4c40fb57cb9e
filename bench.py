#!/usr/bin/env python
"""Benchmark of the STARCOP hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1      # reference CPU path (oracle port)

Metric (BASELINE.json): 512x512 hyperspectral tiles/s for one HyperSTARCOP train step
(forward + weighted BCE + backward + Adam, BatchNorm in training mode).  Workload at every N:
BASELINE.json configs[1] per GPU -- U-Net fwd/bwd on synthetic 512x512x(mag1c+RGB) tiles, bs=16,
bf16 storage / fp32 accumulate (weak scaling: per-GPU work fixed, gradients all-reduced over NCCL).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNET_TRAIN_GFLOP_PER_TILE = 80.24      # fprop + dgrad + wgrad, C=4, 512x512
UNET_FWD_GFLOP_PER_TILE = 26.798


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="tiles per GPU (configs[1]: 16)")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-spectral", action="store_true", help="skip the spectral-product (mag1c / ratio) timings")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the PyTorch-eager-on-GPU baseline legs")
    ap.add_argument("--no-extras", action="store_true", help="skip parity / per-kernel roofline table legs")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of one CUDA graph")
    ap.add_argument("--profile", action="store_true",
                    help="profiling run (under ncu): 1 warm-up, no e2e loop, no roofline probe; prints no bench line")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=20).stdout
                f = [s.strip() for s in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:          # noqa: BLE001
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def cpu_reference_step_rate(size, steps, warmup, tiles=2, stat="median"):
    """Reference CPU path (oracle port of the reference's PyTorch fp32 module, all host threads):
    fwd + loss + bwd + Adam on `tiles` tiles per step; `warmup` untimed + `steps` timed steps."""
    from oracle.module import get_model as oracle_get_model
    from starcop_b200 import synthetic
    from starcop_b200.settings import default_settings
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    m = oracle_get_model(default_settings(pos_weight=1.0))
    opt = m.configure_optimizers()["optimizer"]
    batch = synthetic.hyperstarcop_batch(tiles, size=size, seed=0)
    m.train()
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = m.training_step(batch, i)
        loss.backward()
        opt.step()
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    ts.sort()
    t = ts[len(ts) // 2] if stat == "median" else sum(ts) / len(ts)
    return tiles / t, t, torch.get_num_threads(), f"{tiles} tiles of {size}x{size}x4 per step, {warmup} warm-up + {steps} timed steps, {stat}"


def gpu_eager_baseline(dev, B, S, steps=5, warmup=3):
    """The reference's own GPU path is PyTorch eager (model_module.py:238-251 -> cuDNN): the same module (oracle
    restatement of smp.Unet(mobilenet_v2) + BCE + Adam) on this B200, fp32 as the reference trains (TF32 off and
    on) and bf16 autocast + channels_last -- BASELINE.md B4, SURVEY 8(d) "the real bar".  A reported baseline."""
    from oracle.module import get_model as oracle_get_model
    from oracle import loss_metrics as lm
    from oracle.normalizer import normalize_x
    from starcop_b200 import synthetic
    from starcop_b200.settings import default_settings
    out = {}
    batch = synthetic.hyperstarcop_batch(B, size=S, seed=0)
    x = normalize_x(batch["input"], default_settings().dataset.input_products).to(dev)
    y, w = batch["output"].to(dev), batch["weight_loss"].to(dev)
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    for name, tf32, autocast in (("fp32", False, False), ("fp32_tf32", True, False), ("bf16_autocast_channels_last", True, True)):
        try:
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.benchmark = True
            torch.manual_seed(0)
            m = oracle_get_model(default_settings(pos_weight=1.0)).to(dev).train()
            net = m.network
            xin = x
            if autocast:
                net = net.to(memory_format=torch.channels_last)
                xin = x.contiguous(memory_format=torch.channels_last)
            opt = torch.optim.Adam(net.parameters(), 1e-4, fused=True)

            def step():
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    lg = net(xin)
                loss = torch.mean(lm.bce_with_logits_elementwise(lg.float(), y, m.pos_weight.to(dev)) * w)
                loss.backward()
                opt.step()
                return loss
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"tiles_per_s": B / (ms * 1e-3), "ms_per_step": ms}
            del m, net, opt
            torch.cuda.empty_cache()
        except Exception as e:      # noqa: BLE001
            out[name] = {"error": repr(e)[:200]}
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    out["what"] = (f"PyTorch {torch.__version__} eager, restated smp.Unet(mobilenet_v2) + BCE + fused Adam, bs={B}, {S}x{S}x4, "
                   f"inputs resident, {warmup} warm-up + {steps} timed steps (CUDA events)")
    return out


def mode_parity(dev, S):
    """Distance between the benchmarked bf16 tcgen05 mode and the fp32 parity mode (which tests/test_gpu_parity512.py
    pins to the fp32 oracle at this tile size) on 2 tiles, same weights: eval-mode BatchNorm (what inference /
    batch_with_preds uses) and train-mode BatchNorm.  The bf16 figures equal what PyTorch's own bf16 autocast shows
    against fp32 on this network (same test, profiles/r02_parity512.json): they are the price of bf16 storage, not
    of this implementation; the fp32 mode carries the 1e-4 mask claim."""
    from starcop_b200 import synthetic
    from starcop_b200.model_setup import get_model
    from starcop_b200.settings import default_settings
    b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synthetic.hyperstarcop_batch(2, size=S, seed=5).items()}
    res = {}
    for mode in ("f32", "bf16"):
        torch.manual_seed(1234)
        m = get_model(default_settings(pos_weight=1.0, compute_dtype=mode), None).to(dev).train()
        with torch.no_grad():
            lg = m(b["input"])
            loss = m.training_step(b, 1)
            m.eval()
            le = m(b["input"])
        res[mode] = (torch.sigmoid(lg), float(loss), torch.sigmoid(le))
        del m
    d, de = (res["bf16"][0] - res["f32"][0]).abs(), (res["bf16"][2] - res["f32"][2]).abs()
    return {"what": f"bf16 tcgen05 mode vs fp32 parity mode, 2 tiles {S}x{S}, same weights",
            "eval_bn_sigmoid_max_abs": de.max().item(), "eval_bn_sigmoid_mean_abs": de.mean().item(),
            "train_bn_sigmoid_max_abs": d.max().item(), "train_bn_sigmoid_mean_abs": d.mean().item(),
            "loss_rel": abs(res["bf16"][1] - res["f32"][1]) / abs(res["f32"][1]),
            "fp32_mode_vs_oracle": "eval-mode sigmoid maps <= 1e-4 max-abs (measured 4e-7), loss 1e-5 (tests/test_gpu_parity512.py)",
            "bf16_note": "train-mode BN at initialisation amplifies bf16 rounding (near-constant channels / tiny batch std): "
                         "PyTorch bf16 autocast is equally far from fp32 (tests assert ours <= 1.3x autocast)"}


def step_roofline(model, batch, pk):
    """One eager train step with every kernel launch timed by CUDA events (starcop_b200/profiler.py; weight
    gradients on the main stream for this pass): per-kernel-family table, the time-dominant kernel, and the
    time-weighted aggregates for the whole step and for the convolution stack."""
    from starcop_b200 import profiler
    eng = model.network._engine
    side = eng.side_wgrad
    eng.side_wgrad = False
    try:
        model.train_step_fused(batch)                      # warm
        with profiler.StepProfile() as prof:
            model.train_step_fused(batch)
    finally:
        eng.side_wgrad = side
    tab = prof.table(pk["bf16_tflops_sustained"], pk["hbm_gbs"])
    tot_s = sum(f["s"] for f in tab.values())
    agg = lambda names: (sum(tab[n]["roof_s"] for n in names if n in tab), sum(tab[n]["s"] for n in names if n in tab))
    r_all, s_all = agg(tab.keys())
    r_conv, s_conv = agg(profiler.CONV_STACK)
    r_tc, s_tc = agg(profiler.TENSOR_KERNELS)
    fl_tc = sum(tab[n]["flop"] for n in profiler.TENSOR_KERNELS if n in tab)
    dom = max(tab, key=lambda n: tab[n]["s"])
    d = tab[dom]
    tensor_bound = d["flop"] / (pk["bf16_tflops_sustained"] * 1e12) > d["bytes"] / (pk["hbm_gbs"] * 1e9)
    roof = {"kernel": dom, "launches_per_step": d["launches"], "share_of_step_kernel_time": d["s"] / tot_s,
            "us_per_launch": d["s"] / d["launches"] * 1e6, "how": "CUDA events around every launch of one eager step, in-step cache state"}
    if tensor_bound:
        roof.update({"bound": "tensor", "achieved": d["tflops"], "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": d["tflops"] / pk["bf16_tflops_sustained"], "flop_per_launch": d["flop"] / d["launches"]})
    else:
        roof.update({"bound": "hbm", "achieved": d["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": d["gbs"] / pk["hbm_gbs"], "algorithmic_bytes_per_launch": d["bytes"] / d["launches"]})
    roof["traffic"] = None
    roof["peak_source"] = "MEASURED_PEAKS.json (sustained bf16 / HBM copy): kernel timed inside a long step"
    table = {n: {"launches": f["launches"], "us": round(f["s"] * 1e6, 1), "share": round(f["s"] / tot_s, 4),
                 "tflops": None if not f["flop"] else round(f["tflops"], 1), "gbs": round(f["gbs"], 1),
                 "frac_of_roofline": round(f["frac_of_roofline"], 3)}
             for n, f in sorted(tab.items(), key=lambda kv: -kv[1]["s"])}
    summary = {"sum_kernel_us": round(tot_s * 1e6, 1),
               "step_frac_of_roofline": r_all / s_all,
               "conv_stack_frac_of_roofline": r_conv / s_conv if s_conv else None,
               "conv_stack_share_of_step": s_conv / s_all,
               "tensor_kernels_tflops": fl_tc / s_tc / 1e12 if s_tc else None,
               "tensor_kernels_frac_of_sustained_peak": fl_tc / s_tc / 1e12 / pk["bf16_tflops_sustained"] if s_tc else None,
               "tensor_kernels_frac_of_their_roofline": r_tc / s_tc if s_tc else None,
               "definition": "roofline time per launch = max(FLOP / sustained bf16 peak, algorithmic bytes / HBM peak); fraction = sum(roofline) / sum(measured)"}
    return roof, table, summary


def spectral_products_bench(dev, pk, tiles=8, size=512, bands=125, iters=5):
    """configs[2] front end: 125-band AVIRIS-shape BIP cubes -> mag1c matched filter (+ RGB pick) and the band
    ratio product, timed with CUDA events on resident inputs.  Algorithmic bytes: the whole cube read
    once + 4 output channels written = size*size*(bands+4)*4 per tile; ratio: 12 B / pixel."""
    from starcop_b200 import features, mag1c, synthetic
    g = torch.Generator(device=dev).manual_seed(0)
    c = torch.arange(bands, device=dev, dtype=torch.float32)
    mu = 8.0 * torch.exp(-c / (0.9 * bands)) + 0.6 + 0.15 * torch.sin(c * 0.37)
    albedo = torch.nn.functional.interpolate(torch.rand(tiles, 1, 9, 9, device=dev, generator=g) + 0.5, size=(size, size),
                                             mode="bilinear", align_corners=True)[:, 0]
    cube = albedo[..., None] * mu * (1.0 + 0.01 * torch.randn(tiles, size, size, bands, device=dev, generator=g))
    t = torch.as_tensor(synthetic.synthetic_template(73), device=dev)
    cube[:, 200:260, 100:180, 52:] *= (1.0 + 0.02 * t.float())
    sl = slice(52, 125)
    tmpl = synthetic.synthetic_template(73)

    def timeit(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e-3

    out = {}
    tile_bytes = size * size * (bands + 4) * 4
    for name, it in (("mag1c_rmf", 0), ("mag1c_acrwl1mf_30it", 30)):
        sec = timeit(lambda: mag1c.mag1c_tiles(cube, tmpl, sl, num_iter=it))
        gbs = tiles * tile_bytes / sec / 1e9
        out[name] = {"tiles_per_s": tiles / sec, "us_per_tile": sec / tiles * 1e6, "achieved_gbs": gbs,
                     "frac_of_hbm_peak": gbs / pk["hbm_gbs"], "bound": "per-group S x S statistics, not HBM: exact covariance by tcgen05 digit-slice MMAs, 73 dependent fp64 sweep pivots, "
                                                                              "two shared-memory passes per reweighting iteration (mag1c_tc.cu)"}
    bg, sig = cube[..., 100].contiguous(), cube[..., 110].contiguous()
    sec = timeit(lambda: features.ratio_2c_match_c_from_sums_outlier(bg, sig))
    gbs = tiles * size * size * 12 / sec / 1e9
    out["ratio_2c_outlier"] = {"tiles_per_s": tiles / sec, "us_per_tile": sec / tiles * 1e6, "achieved_gbs": gbs,
                               "frac_of_hbm_peak": gbs / pk["hbm_gbs"], "bound": "latency of the exact radix select (cluster barriers + histogram scans); HBM traffic is the algorithmic 12 B / px"}
    # SRF band aggregation (aviris.py:262-338): the single-pass per-pixel spectral product -- every cube byte read once,
    # 8 simulated bands written: algorithmic bytes = size*size*(bands + 8)*4 per tile
    from starcop_b200 import srf
    centers = 380.0 + 17.0 * np.arange(bands)
    wl = np.arange(400.0, 2400.0, 2.0)
    resp = np.stack([np.exp(-0.5 * ((wl - (450 + 230 * k)) / 40.0) ** 2) for k in range(8)])
    Wt = srf.srf_weight_table(wl, resp, centers)
    sec = timeit(lambda: srf.transform_to_srf(cube, Wt, 0.0))
    gbs = tiles * size * size * (bands + 8) * 4 / sec / 1e9
    out["srf_aggregate_8_bands"] = {"tiles_per_s": tiles / sec, "us_per_tile": sec / tiles * 1e6, "achieved_gbs": gbs,
                                    "frac_of_hbm_peak": gbs / pk["hbm_gbs"], "bound": "hbm (one pass: cube read once, 8 bands written)"}
    out["workload"] = f"{tiles} cubes of {size}x{size}x{bands} f32 BIP, 73-band SWIR window, groups = detector columns"
    return out


def chain_bench(dev, pk, dtype, B=8, size=512, bands=125, steps=5):
    """BASELINE.json configs[2]: bs=8 cubes of 512x512x125 -> mag1c (30 iterations, groups = detector columns) -> fused
    pack (RGB pick + normalise + NHWC, weight_mag1c) -> U-Net train step (fwd + BCE + bwd + Adam), one timed chain
    on resident cubes (4 x 131 MB cubes rotate so that no step finds its cube in L2)."""
    from starcop_b200 import chain, synthetic
    from starcop_b200.model_setup import get_model
    from starcop_b200.settings import default_settings
    torch.manual_seed(0)
    model = get_model(default_settings(pos_weight=1.0, compute_dtype=dtype), None).to(dev).train()
    g = torch.Generator(device=dev).manual_seed(0)
    c = torch.arange(bands, device=dev, dtype=torch.float32)
    mu = 8.0 * torch.exp(-c / (0.9 * bands)) + 0.6 + 0.15 * torch.sin(c * 0.37)
    cubes = []
    for _ in range(2):
        albedo = torch.nn.functional.interpolate(torch.rand(B, 1, 9, 9, device=dev, generator=g) + 0.5, size=(size, size),
                                                 mode="bilinear", align_corners=True)[:, 0]
        cubes.append(albedo[..., None] * mu * (1.0 + 0.01 * torch.randn(B, size, size, bands, device=dev, generator=g)))
    tmpl = synthetic.synthetic_template(73)
    sl = slice(52, 125)
    rgb = (40, 25, 10)
    y = (torch.rand(B, 1, size, size, device=dev, generator=g) > 0.98).float()

    def step(i):
        b = chain.cube_batch(model, cubes[i % 2], tmpl, sl, rgb, output=y)
        return model.train_step_fused(b)
    for i in range(2):
        step(i)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    for i in range(steps):                                  # the filter stage alone, same cubes
        chain.cube_batch(model, cubes[i % 2], tmpl, sl, rgb, output=y)
    e2.record()
    torch.cuda.synchronize()
    ms, ms_f = e0.elapsed_time(e1) / steps, e1.elapsed_time(e2) / steps
    cube_bytes = B * size * size * (bands + 4) * 4
    return {"workload": f"configs[2]: {B} cubes of {size}x{size}x{bands} f32 BIP -> mag1c (acrwl1mf, 30 iterations) -> fused pack -> "
                        f"U-Net train step, {dtype}", "tiles_per_s": B / (ms * 1e-3), "ms_per_step": ms,
            "ms_spectral_stage": ms_f, "ms_unet_stage": ms - ms_f,
            "spectral_stage_gbs": cube_bytes / (ms_f * 1e-3) / 1e9, "spectral_stage_frac_of_hbm_peak": cube_bytes / (ms_f * 1e-3) / 1e9 / pk["hbm_gbs"],
            "steps": steps}


def emit_sweep_bench(dev, pk, dtype, rows=1280, cols=1242, bands=285):
    """BASELINE.json configs[4] on one GPU: an EMIT-shape granule (rows x cols x 285 f32 BIP) -> mag1c_emit (fp64,
    621 two-column groups, alpha 1e-4, 30 iterations; mag1c_emit.py:16-90) -> RGB pick + EMIT rescale -> sigmoid(U-Net)
    in eval mode, un-tiled (padded_predict, the notebook's call) and tiled at 256 / 512 / 1024 px."""
    from starcop_b200 import emit, synthetic
    from starcop_b200.model_setup import get_model
    from starcop_b200.settings import default_settings
    torch.manual_seed(0)
    model = get_model(default_settings(pos_weight=1.0, compute_dtype=dtype), None).to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(1)
    wl = np.linspace(381.0, 2493.0, bands)
    c = torch.arange(bands, device=dev, dtype=torch.float32)
    mu = 8.0 * torch.exp(-c / (0.9 * bands)) + 0.6
    raw = (torch.rand(rows, cols, 1, device=dev, generator=g) + 0.5) * mu * (1.0 + 0.01 * torch.randn(rows, cols, bands, device=dev, generator=g))
    S = int(((wl >= 2122) & (wl <= 2488)).sum())
    tmpl = synthetic.synthetic_template(S)

    def timeit(fn, n=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    out = {"workload": f"configs[4] (one GPU): EMIT-shape granule {rows}x{cols}x{bands} f32 BIP, {S}-band window, fp64 matched filter in "
                       f"{(cols + 1) // 2} two-column groups, eval-mode {dtype} U-Net"}
    x = {}
    ms_f = timeit(lambda: x.update(inp=emit.emit_model_input(raw, wl, template=tmpl, column_step=2)[0]), n=2)
    cube_bytes = rows * cols * (bands + 4) * 4
    out["mag1c_emit_fp64"] = {"ms": ms_f, "groups_per_s": ((cols + 1) // 2) / (ms_f * 1e-3), "cube_gbs": cube_bytes / (ms_f * 1e-3) / 1e9,
                              "frac_of_hbm_peak": cube_bytes / (ms_f * 1e-3) / 1e9 / pk["hbm_gbs"],
                              "reference": "28 s on Colab CPU (21.8 groups/s), notebooks/inference_on_raw_EMIT_nc_file.ipynb:282"}
    scene = x["inp"]
    px = scene.shape[-1] * scene.shape[-2]
    for tile in (None, 256, 512, 1024):
        with torch.no_grad():
            ms = timeit(lambda: emit.predict_scene(scene, model, tile=tile, batch=16))
        out["unet_" + ("untiled" if tile is None else f"tile{tile}")] = {"ms": ms, "mpx_per_s": px / (ms * 1e-3) / 1e6,
                                                                        "tiles512_equiv_per_s": px / (512 * 512) / (ms * 1e-3),
                                                                        "fwd_tflops": px / (512 * 512) * UNET_FWD_GFLOP_PER_TILE / ms}
    return out


def run_reference(args, rank):
    """--impl reference: the reference's own CPU path (oracle port of its PyTorch fp32 module; the reference itself
    cannot be installed here) on all host threads, on this arm's config: `--batch` tiles of
    size x size x 4 per step, EXACTLY `--warmup` untimed and `--steps` timed steps (mean step time)."""
    if rank != 0:
        return
    rate, mean_s, cores, sample = cpu_reference_step_rate(args.size, args.steps, args.warmup, tiles=args.batch, stat="mean")
    line = {"impl": "reference", "metric": "512x512 hyperspectral tiles/sec (train step)", "value": rate,
            "unit": "tiles/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[1] on the host CPU: HyperSTARCOP U-Net fwd+bwd+Adam, {args.size}x{args.size}x(mag1c+RGB) "
                                   f"tiles, bs={args.batch}, PyTorch fp32, BN train mode",
                       "global_batch": args.batch, "tile": [args.size, args.size, 4]},
            "cpu_baseline": {"value": rate, "unit": "tiles/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def emit(line):
    """the ONE JSON line goes to the real stdout; everything else this process (or a library under it: the NCCL
    version banner, torch warnings) prints is routed to stderr"""
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from starcop_b200 import _lib, synthetic
    from starcop_b200.model_setup import get_model
    from starcop_b200.settings import default_settings

    from starcop_b200 import parallel
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    parallel.init_distributed("nccl")
    _lib.load()

    torch.manual_seed(0)                       # same initial weights on every rank (DDP broadcast equivalent)
    model = get_model(default_settings(pos_weight=1.0, compute_dtype=args.dtype), None).to(dev)
    model.train()
    B, S = args.batch, args.size
    host = synthetic.hyperstarcop_batch(B, size=S, seed=100 + rank)        # rank-seeded tiles
    pinned = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host.items()}
    resident = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}

    parallel.broadcast_parameters(model.network.flat_params, [b for _, b in model.network.named_buffers()])
    grad_sync = parallel.GradSync(world)       # NCCL sum over NVLink; the mean is folded into Adam's grad_scale

    # the whole step (forward, loss, backward, all-reduce, Adam) replayed as one CUDA graph
    graphed = None if args.no_graph else model.make_graphed_train_step(resident, grad_sync=grad_sync, double_buffer=True)

    def step_resident():
        if graphed is not None:
            return graphed(resident)
        return model.train_step_fused(resident, grad_sync=grad_sync)

    staging = {k: torch.empty_like(v, device=dev) for k, v in host.items() if torch.is_tensor(v)}
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    e2e_state = {"i": 0, "n": 0}

    def step_e2e():
        if graphed is not None:
            # every step's batch is copied from pinned host memory inside the timed region; the copy of step
            # i+1 runs on the copy stream while step i computes (pinned DataLoader + non_blocking transfer)
            if e2e_state["i"] == 0:
                graphed.prefetch(pinned)
            e2e_state["i"] += 1
            if e2e_state["i"] < e2e_state["n"]:
                graphed.prefetch(pinned)
            loss = graphed()
        else:
            for k, v in pinned.items():
                if torch.is_tensor(v):
                    staging[k].copy_(v, non_blocking=True)
            loss = model.train_step_fused(staging, grad_sync=grad_sync)
        loss_host.copy_(loss.reshape(1), non_blocking=True)
        return loss

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = _lib.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, _lib.launch_count() - c0

    if args.profile:
        timed(step_resident, args.steps, 1)
        return
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches_per_step = None
    if graphed is not None:                     # kernel launches inside the graph = C-ABI calls of one eager step
        c0 = _lib.launch_count()
        model.train_step_fused(resident, grad_sync=grad_sync)
        launches_per_step = _lib.launch_count() - c0
    ms, launches = timed(step_resident, args.steps, max(args.warmup, 3))
    if launches_per_step is not None:
        launches = launches_per_step * args.steps
    def e2e_run(n):
        e2e_state["i"], e2e_state["n"] = 0, n
        return n
    # warm-up and timed loop are separate prefetch chains (each copies exactly one batch per step)
    e2e_run(1)
    step_e2e()
    torch.cuda.synchronize()
    e2e_run(args.steps)
    ms_e2e, _ = timed(step_e2e, args.steps, 0)
    collective = None
    if world > 1 and graphed is not None:
        # exposed time of the data-parallel exchange: the same graphed step captured WITHOUT the gradient all-reduce
        # (weights then drift apart across ranks, which is why this leg runs last), max over ranks like every timing
        nosync = model.make_graphed_train_step(resident, grad_sync=None, warmup=1)
        ms_ns, _ = timed(lambda: nosync(resident), args.steps, 3)
        nbytes = model.network.flat_grads.numel() * 4
        split = model.network.decoder_grad_offset()
        collective = {"exposed_ms_per_step": (ms - ms_ns) / args.steps, "ms_per_step_without_exchange": ms_ns / args.steps,
                      "bytes_per_step": nbytes,
                      "buckets": ([{"what": "decoder + head, all-reduce launched when their weight gradients are issued (runs under "
                                            "the encoder's backward)", "bytes": nbytes - 4 * split},
                                   {"what": "encoder, after the backward pass", "bytes": 4 * split}]
                                  if getattr(grad_sync, "bucketed", False) else [{"what": "whole arena after the backward pass", "bytes": nbytes}]),
                      "backend": "NCCL all-reduce (sum) inside the step's CUDA graph; 1/world folded into Adam"}
        nosync = None
    sampler.stop_flag = True

    tiles = world * B * args.steps
    value = tiles / (ms / 1e3)
    e2e_val = tiles / (ms_e2e / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in host.values() if torch.is_tensor(v))

    pk, pk_kind = peaks()
    pk.setdefault("bf16_tflops_sustained", pk.get("bf16_tflops", 1400.0))
    roof = table = summary = None
    if rank == 0 and world == 1 and not args.no_extras:
        roof, table, summary = step_roofline(model, resident, pk)
        roof["peak_source"] = f"{pk_kind}: " + roof["peak_source"]
    probe = model.network.bench_dominant_kernel(B, S) if hasattr(model.network, "bench_dominant_kernel") else None
    if probe is not None:
        probe["peak"] = pk["bf16_tflops"]
        probe["frac"] = probe["achieved"] / probe["peak"]
        probe["peak_source"] = f"{pk_kind} burst bf16 (kernel timed alone, cold rotating buffers)"

    spectral = None
    if rank == 0 and world == 1 and not args.no_spectral:
        try:
            spectral = spectral_products_bench(dev, pk)
        except Exception as e:          # noqa: BLE001  (reported, never silently dropped)
            spectral = {"error": repr(e)}
    chain_res = emit_res = None
    if rank == 0 and world == 1 and not args.no_spectral:
        for name, fn in (("chain", lambda: chain_bench(dev, pk, args.dtype)), ("emit", lambda: emit_sweep_bench(dev, pk, args.dtype))):
            try:
                r = fn()
            except Exception as e:          # noqa: BLE001
                r = {"error": repr(e)[:300]}
            if name == "chain":
                chain_res = r
            else:
                emit_res = r
            torch.cuda.empty_cache()
    eager = parity = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            parity = mode_parity(dev, S)
        except Exception as e:          # noqa: BLE001
            parity = {"error": repr(e)[:200]}
    if rank == 0 and world == 1 and not args.no_eager_baseline:
        del graphed
        torch.cuda.empty_cache()
        eager = gpu_eager_baseline(dev, B, S)
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            rate, med, cores, sample = cpu_reference_step_rate(S, 3, 1)
            cpu = {"value": rate, "unit": "tiles/s", "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": "512x512 hyperspectral tiles/sec (train step)", "value": value, "unit": "tiles/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"configs[1]: HyperSTARCOP U-Net fwd+bwd+Adam, {S}x{S}x(mag1c+RGB) tiles, bs={B}/GPU, "
                                   f"{args.dtype} storage fp32 accumulate, BN train mode",
                       "global_batch": world * B, "tile": [S, S, 4], "parallelism": f"dp{world}",
                       "l2": "per-step working set (activations > 2 GB) exceeds the 126 MB L2",
                       "schedule": "one CUDA graph per step captured on a " + ("default" if os.environ.get("STARCOP_MAIN_PRIO") == "0" else "high")
                                   + "-priority stream; weight gradients on "
                                   + os.environ.get("STARCOP_SIDE_STREAMS", "2") + " default-priority side stream(s); "
                                   "programmatic dependent launch " + ("off" if os.environ.get("STARCOP_PDL") == "0" else "on")},
            "e2e": {"value": e2e_val, "unit": "tiles/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "unet_tflops": value * UNET_TRAIN_GFLOP_PER_TILE / 1e3,
            "unet_frac_of_sustained_bf16_peak": value * UNET_TRAIN_GFLOP_PER_TILE / 1e3 / pk["bf16_tflops_sustained"],
            "roofline": roof, "roofline_summary": summary, "roofline_table": table,
            "largest_gemm_probe": probe,
            "cpu_baseline": cpu, "gpu_eager_baseline": eager, "parity": parity, "spectral_products": spectral,
            "configs2_cube_to_unet_chain": chain_res, "configs4_emit_inference_sweep": emit_res,
            "collective": collective,
        }
        emit(line)
    if world > 1:
        # Orderly teardown: the captured CUDA graphs hold NCCL work, so they are destroyed BEFORE the communicator
        # (destroying the process group under live graphs was observed to hang in round 1).  A watchdog ends the
        # process if the teardown still stalls: the result line is already out, a hang must never eat the run.
        import gc
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        watchdog = threading.Timer(30.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        graphed = None
        model = None
        gc.collect()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        watchdog.cancel()


if __name__ == "__main__":
    main()
