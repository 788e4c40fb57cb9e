"""Whole-network parity: the CUDA HyperSTARCOP U-Net / ModelModule against the CPU oracle (which
is pinned to the reference by tests/test_oracle_golden.py) and against the golden vectors the
reference itself produced."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.module import get_model as oracle_get_model  # noqa: E402
from starcop_b200 import synthetic  # noqa: E402
from starcop_b200.model_setup import get_model  # noqa: E402
from starcop_b200.settings import default_settings  # noqa: E402

DEV = "cuda"


def to_dev(batch):
    return {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in batch.items()}


def build_pair(pw=1.0, compute_dtype="f32", seed=1234):
    torch.manual_seed(seed)
    oracle = oracle_get_model(default_settings(pos_weight=pw))
    torch.manual_seed(seed)
    model = get_model(default_settings(pos_weight=pw, compute_dtype=compute_dtype), None).to(DEV)
    return oracle, model


def rel_err(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def f64_truth(pw, batch):
    """The same network evaluated in float64 on the CPU: the yardstick for fp32 rounding noise.
    (Conv weights that feed a BatchNorm have near-zero true gradients -- the loss is invariant to
    their scale -- so fp32 noise on them is ~1e-2 of max|grad| for PyTorch's own fp32 path too.)"""
    from oracle import loss_metrics as lm
    from oracle.normalizer import normalize_x
    torch.manual_seed(1234)
    o = oracle_get_model(default_settings(pos_weight=pw)).train().double()
    x = normalize_x(batch["input"], o.input_products).double()
    loss = torch.mean(lm.bce_with_logits_elementwise(o.network(x), batch["output"].double(), o.pos_weight) *
                      batch["weight_loss"].double())
    loss.backward()
    return loss.item(), {n: p.grad for n, p in o.network.named_parameters()}


@pytest.mark.parametrize("pw", [1.0, 15.0])
def test_train_step_fp32_matches_oracle_and_reference(golden, pw):
    g = golden("model_module.npz")
    oracle, model = build_pair(pw)
    batch = synthetic.hyperstarcop_batch(2, size=64, seed=3)
    oracle.train(); model.train()
    lo = oracle.training_step(batch, 0); lo.backward()
    lm_ = model.training_step(to_dev(batch), 0); lm_.backward()
    l64, g64 = f64_truth(pw, batch)
    tag = f"pw{int(pw)}"
    # loss: fp32 tolerance 1e-5 relative (reference golden, fp32 oracle, fp64 truth)
    assert abs(lm_.item() - float(g[f"{tag}_train_loss"])) <= 1e-5 * abs(float(g[f"{tag}_train_loss"]))
    assert abs(lm_.item() - lo.item()) <= 1e-5 * abs(lo.item())
    assert abs(lm_.item() - l64) <= 1e-5 * abs(l64)
    # every parameter gradient: the CUDA path must be as close to the float64 truth as the reference's own
    # fp32 PyTorch arithmetic is.  Both fp32 evaluations are NOISE around the float64 value (at this tile size
    # the deep layers normalise over 8...32 pixels, and one ReLU/ReLU6 mask flipped by a 1e-7 perturbation moves
    # a channel's gradient by percents -- for PyTorch's CPU kernels too), so the comparison is statistical:
    # typical error no worse than the reference's (median ratio <= 1.5), >= 95 % of the parameters within
    # 3x its error (floor 1e-4), every parameter within 10x (floor 1e-3).
    ratios, worst = [], []
    for (n, po), (_, pm) in zip(oracle.network.named_parameters(), model.network.named_parameters()):
        assert pm.grad is not None, n
        ref = g64[n]
        e_gpu = rel_err(pm.grad.cpu().double(), ref)
        e_cpu = rel_err(po.grad.double(), ref)
        assert e_gpu <= max(10 * e_cpu, 1e-3), (n, e_gpu, e_cpu)
        ratios.append(e_gpu / max(e_cpu, 1e-4 / 3))
        if e_gpu > max(3 * e_cpu, 1e-4):
            worst.append((n, e_gpu, e_cpu))
    assert len(worst) <= 0.05 * len(ratios), worst
    assert float(np.median(ratios)) <= 1.5, float(np.median(ratios))
    # well-conditioned gradients (head, last decoder block) agree tightly with the reference's vectors
    assert np.allclose(model.network.segmentation_head[0].weight.grad.cpu().numpy(), g[f"{tag}_grad_head_w"], rtol=1e-3, atol=1e-6)
    e = rel_err(model.network.decoder.blocks[4].conv2[1].weight.grad.cpu(), torch.from_numpy(g[f"{tag}_grad_dec_b4c2_bn_w"]))
    assert e <= 1e-3, e
    # BN running statistics updated like torch
    for (n, bo), (_, bm) in zip(oracle.network.named_buffers(), model.network.named_buffers()):
        if n.endswith("num_batches_tracked"):
            assert int(bo) == int(bm) == 1
        else:
            assert torch.allclose(bm.cpu(), bo, rtol=1e-4, atol=1e-5), n


def test_eval_forward_and_batch_with_preds(golden):
    g = golden("model_module.npz")
    oracle, model = build_pair(1.0)
    batch = synthetic.hyperstarcop_batch(2, size=64, seed=3)
    # one train forward first so running stats are non-trivial on both sides
    oracle.train(); model.train()
    with torch.no_grad():
        oracle.training_step(batch, 0)
        model.training_step(to_dev(batch), 0)
    oracle.eval(); model.eval()
    with torch.no_grad():
        bo = oracle.batch_with_preds(batch)
        bm = model.batch_with_preds(to_dev(batch))
    # segmentation masks (sigmoid maps) within 1e-4 max-abs of the reference path
    assert (bm["prediction"].cpu() - bo["prediction"]).abs().max().item() <= 1e-4
    assert (bm["logits"].cpu() - bo["logits"]).abs().max().item() <= 1e-3
    assert torch.equal(bm["input_norm"].cpu(), bo["input_norm"])
    # integer outputs: bit exact wherever the oracle logit is not within the fp32 tolerance of 0
    safe = bo["logits"].abs() > 1e-3
    for k in ("pred_binary", "differences"):
        assert torch.equal(bm[k].cpu()[safe], bo[k][safe]), k
    assert torch.equal(bm["pred_classification"].cpu(), bo["pred_classification"])
    assert torch.allclose(bm["loss_per_pixel"].cpu(), bo["loss_per_pixel"], rtol=1e-4, atol=1e-4)
    assert torch.allclose(bm["loss_per_pixel_weighted"].cpu(), bo["loss_per_pixel_weighted"], rtol=1e-4, atol=1e-4)


def test_val_step_confusion_matrices(golden):
    g = golden("model_module.npz")
    oracle, model = build_pair(1.0)
    batch = synthetic.hyperstarcop_batch(2, size=64, seed=3)
    oracle.train(); model.train()
    oracle.training_step(batch, 0).backward()      # the golden vectors were taken after one train forward
    model.training_step(to_dev(batch), 0)
    oracle.eval(); model.eval()
    with torch.no_grad():
        lo = oracle.val_step(batch, 0)
        lv = model.val_step(to_dev(batch), 0)
    assert abs(lv.item() - lo.item()) <= 1e-5 * abs(lo.item())
    assert torch.equal(model.confusion_matrix.compute().cpu(), oracle.cm)
    assert torch.equal(model.classification_confusion_matrix.compute().cpu(), oracle.cm_cls)
    assert np.array_equal(model.confusion_matrix.compute().cpu().numpy(), g["pw1_val_cm"])
    assert np.array_equal(model.classification_confusion_matrix.compute().cpu().numpy(), g["pw1_val_cm_cls"])
    model.val_epoch_end(None, "val")
    assert "val_iou" in model._logged and "val_classification_f1score" in model._logged
    assert int(model.confusion_matrix.compute().sum()) == 0


def test_state_dict_layout_is_reference_compatible(golden):
    g = golden("model_module.npz")
    _, model = build_pair(1.0)
    assert list(model.state_dict().keys()) == [str(k) for k in g["state_dict_keys"]] or \
        set(model.state_dict().keys()) == set(str(k) for k in g["state_dict_keys"])
    sums = {k: float(v.double().sum()) for k, v in model.state_dict().items()}
    ref = {str(k): float(s) for k, s in zip(g["state_dict_keys"], g["state_dict_sums"])}
    for k, s in ref.items():
        assert abs(sums[k] - s) <= 1e-9 * max(1.0, abs(s)), k     # same seed -> same initial weights


def test_fused_train_steps_track_oracle_adam():
    """5 optimisation steps: fused CUDA step (fwd+loss+bwd+Adam) vs oracle + torch.optim.Adam."""
    oracle, model = build_pair(1.0)
    opt = oracle.configure_optimizers()["optimizer"]
    oracle.train(); model.train()
    for step in range(5):
        batch = synthetic.hyperstarcop_batch(2, size=64, seed=10 + step)
        opt.zero_grad()
        lo = oracle.training_step(batch, step); lo.backward(); opt.step()
        lf = model.train_step_fused(to_dev(batch))
        # Adam's first updates are ~lr*sign(g): parameters whose gradient is fp32 noise move in
        # rounding-dependent directions, so trajectories agree to ~1e-3, not to fp32 epsilon
        assert abs(lf.item() - lo.item()) <= (1e-5 if step == 0 else 1e-2) * abs(lo.item()), (step, lf.item(), lo.item())


def test_bf16_mode_tracks_fp32_within_stated_tolerance():
    """bf16 storage / fp32 accumulate (BASELINE.json's throughput mode).  Stated tolerance: loss
    within 2e-2 relative of the fp32 oracle, logits within 5e-2 of max|logit|; gradients are judged
    against PyTorch's own bf16 autocast of the same network (same rounding points: bf16 conv
    in/out, fp32 BatchNorm arithmetic), because bf16 noise on the ill-conditioned pre-BN weight
    gradients is inherent, not an implementation property."""
    oracle, model = build_pair(1.0, compute_dtype="bf16")
    batch = synthetic.hyperstarcop_batch(2, size=64, seed=3)
    oracle.train(); model.train()
    lo = oracle.training_step(batch, 0)
    lm_ = model.training_step(to_dev(batch), 0)
    lm_.backward()
    assert abs(lm_.item() - lo.item()) <= 2e-2 * abs(lo.item())
    lo.backward()
    # PyTorch bf16 autocast of the oracle network on the GPU
    torch.manual_seed(1234)
    auto = oracle_get_model(default_settings(pos_weight=1.0)).to(DEV).train()
    from oracle import loss_metrics as lmx
    from oracle.normalizer import normalize_x
    xb = normalize_x(batch["input"], auto.input_products).to(DEV)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        lg = auto.network(xb)
    la = torch.mean(lmx.bce_with_logits_elementwise(lg.float(), batch["output"].to(DEV), auto.pos_weight) *
                    batch["weight_loss"].to(DEV))
    la.backward()
    cos = lambda a, b: torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()
    for name in ("segmentation_head.0.weight", "decoder.blocks.4.conv2.0.weight", "decoder.blocks.4.conv1.0.weight",
                 "decoder.blocks.3.conv2.0.weight"):
        go = dict(oracle.network.named_parameters())[name].grad
        c_mine = cos(dict(model.network.named_parameters())[name].grad.cpu(), go)
        c_auto = cos(dict(auto.network.named_parameters())[name].grad.cpu(), go)
        assert c_mine >= c_auto - 0.1, (name, c_mine, c_auto)
    assert cos(model.network.segmentation_head[0].weight.grad.cpu(), oracle.network.segmentation_head[0].weight.grad) > 0.999


def test_256_pixel_tiles_keep_the_deepest_pointwise_convs_on_tcgen05(monkeypatch):
    """BASELINE.json configs[4] sweeps 256-pixel tiles: their 1/32-resolution maps are 8 x 8, which do not tile into the
    tensor-core kernel's 8 x 16 patches.  1 x 1 convolutions are per-pixel operations, so the engine presents the same
    pixels as N*64/128 images of 8 x 16 (engine._tc_dims): no fp32-FMA fallback launches, and the step equals the
    128-pixel-wide case of the same pixels (pointwise kernels do not care which image a pixel belongs to)."""
    from starcop_b200 import engine, profiler
    model = get_model(default_settings(pos_weight=1.0, compute_dtype="bf16"), None).to(DEV).train()
    batch = to_dev(synthetic.hyperstarcop_batch(4, size=256, seed=7))
    names = []
    inner = engine.call
    monkeypatch.setattr(engine, "call", lambda name, *a: (names.append(name), inner(name, *a))[1])
    l1 = model.train_step_fused(batch)
    torch.cuda.synchronize()
    assert "sc_conv_fprop" not in names and "sc_conv_wgrad" not in names, sorted(set(names))
    assert names.count("sc_tc_conv_fprop") > 60
    # reference for the regrouping: the engine's own fp32-FMA kernels on the 8 x 8 maps (no regrouping there)
    monkeypatch.setattr(engine.UNetEngine, "_tc_dims", staticmethod(lambda x, k, stride: (x.N, x.H, x.W)))
    torch.manual_seed(1234)
    ref = get_model(default_settings(pos_weight=1.0, compute_dtype="bf16"), None).to(DEV).train()
    torch.manual_seed(1234)
    mod = get_model(default_settings(pos_weight=1.0, compute_dtype="bf16"), None).to(DEV).train()
    names.clear()
    lr_ = ref.train_step_fused(batch)
    assert "sc_conv_fprop" in names                      # 8 x 8 maps fall back without the regrouping
    monkeypatch.undo()
    lm_ = mod.train_step_fused(batch)
    assert abs(lm_.item() - lr_.item()) <= 2e-3 * abs(lr_.item())
    d = (mod.network.flat_params - ref.network.flat_params).abs().max().item()
    assert d <= 2.5e-3, d                                # one Adam step of lr 1e-3: every weight moved by <= ~1e-3


def test_requires_cuda_and_divisible_by_32():
    _, model = build_pair(1.0)
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 4, 48, 64, device=DEV))
    with pytest.raises(Exception):
        model(torch.zeros(1, 4, 64, 64))           # CPU tensor: no CPU path


@pytest.mark.parametrize("compute_dtype", ["f32", "bf16"])
def test_train_step_is_bit_reproducible(compute_dtype):
    """No floating-point atomics anywhere in the step (weight gradients: per-split partial tiles summed in split
    order; BatchNorm: partial rows; loss: block partials summed in block order): two runs from the same weights on
    the same batch give bit-identical gradients, losses and updated parameters."""
    outs = []
    for _ in range(2):
        _, m = build_pair(1.0, compute_dtype=compute_dtype)
        m.train()
        losses = []
        for s_ in range(2):
            b = to_dev(synthetic.hyperstarcop_batch(2, size=96, seed=30 + s_))
            losses.append(m.train_step_fused(b))
        outs.append((torch.stack(losses).clone(), m.network.flat_grads.clone(), m.network.flat_params.clone()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("compute_dtype", ["f32", "bf16"])
def test_graphed_train_step_matches_eager(compute_dtype):
    """The CUDA-graph replay of the fused step is the same computation as launching it eagerly: after three
    optimisation steps the two models hold BIT-IDENTICAL parameters (the step is deterministic)."""
    _, m1 = build_pair(1.0, compute_dtype=compute_dtype)
    _, m2 = build_pair(1.0, compute_dtype=compute_dtype)
    m1.train(); m2.train()
    b0 = to_dev(synthetic.hyperstarcop_batch(2, size=64, seed=20))
    m1.train_step_fused(b0)                                  # the graph builder runs one eager warm-up step
    step = m2.make_graphed_train_step(b0, warmup=1)
    for s in range(3):
        b = to_dev(synthetic.hyperstarcop_batch(2, size=64, seed=20 + s))
        l1 = m1.train_step_fused(b)
        l2 = step(b)
        assert l1.item() == l2.item(), (s, l1.item(), l2.item())
    assert torch.equal(m1.network.flat_params, m2.network.flat_params)
    for (n, b1), (_, b2) in zip(m1.network.named_buffers(), m2.network.named_buffers()):
        assert torch.equal(b1, b2), n
    assert int(m2.network._adam_state["step"].item()) == 4   # one eager warm-up + 3 replays (capture records, it does not execute)


def test_lr_scheduler_drives_the_fused_step():
    """ReduceLROnPlateau (model_module.py:178-185) on the fused / graphed step: the scheduler's learning rate is
    pushed to the device-resident lr with set_lr, the captured graph picks it up without re-capture."""
    _, m = build_pair(1.0)
    m.train()
    b = to_dev(synthetic.hyperstarcop_batch(2, size=64, seed=21))
    step = m.make_graphed_train_step(b, warmup=1)
    cfg = m.configure_optimizers()
    sched, opt = cfg["lr_scheduler"], cfg["optimizer"]
    assert isinstance(sched, torch.optim.lr_scheduler.ReduceLROnPlateau)
    step(b)
    p0 = m.network.flat_params.clone()
    for _ in range(m.lr_patience + 2):                      # a flat val_loss: the plateau scheduler halves the lr once
        sched.step(1.0)
    new_lr = opt.param_groups[0]["lr"]
    assert abs(new_lr - m.lr * m.lr_decay) <= 1e-12
    m.network.set_lr(new_lr)
    step(b)
    d_small = (m.network.flat_params - p0).abs().max().item()
    m.network.set_lr(0.0)
    p1 = m.network.flat_params.clone()
    step(b)
    assert torch.equal(m.network.flat_params, p1)           # lr = 0: the replayed graph reads the device lr
    assert 0 < d_small <= 4 * new_lr                        # |Adam update| <= lr * (1-b1)/sqrt(1-b2) per element


def test_eval_forward_between_forward_and_backward_keeps_the_tape():
    """an eval / no_grad forward (logging, batch_with_preds) between training_step and .backward() must
    not clobber the pending backward's activations; a second RECORDING forward must make the stale backward fail
    loudly instead of computing garbage."""
    _, m1 = build_pair(1.0)
    _, m2 = build_pair(1.0)
    m1.train(); m2.train()
    b = to_dev(synthetic.hyperstarcop_batch(2, size=64, seed=22))
    other = to_dev(synthetic.hyperstarcop_batch(2, size=64, seed=23))
    l1 = m1.training_step(b, 0); l1.backward()
    l2 = m2.training_step(b, 0)
    with torch.no_grad():
        m2(other["input"])                                   # train-mode no_grad forward
    m2.eval()
    m2.batch_with_preds(other)                               # eval forward
    out = m2(other["input"])                                 # eval forward WITHOUT no_grad (reference allows it)
    assert not out.requires_grad
    m2.train()
    l2.backward()
    for (n, p1), (_, p2) in zip(m1.network.named_parameters(), m2.network.named_parameters()):
        assert torch.equal(p1.grad, p2.grad), n
    la = m2.training_step(b, 0)
    lb = m2.training_step(other, 0)
    with pytest.raises(Exception, match="overwritten"):
        la.backward()
    lb.backward()


def test_bf16_size_change_after_plan_complete():
    """layers that become tensor-core eligible only at a larger input size (the 1/16 and 1/32
    resolution layers fail W % 16 at 128 x 128) get their bf16 weight packing the first time they are requested,
    even though the step's batched re-pack has already run."""
    _, mb = build_pair(1.0, compute_dtype="bf16")
    _, mf = build_pair(1.0, compute_dtype="f32")
    small = to_dev(synthetic.hyperstarcop_batch(2, size=128, seed=24))
    big = to_dev(synthetic.hyperstarcop_batch(2, size=512, seed=25))
    mb.train(); mf.train()
    for m in (mb, mf):
        m.train_step_fused(small)
        m.train_step_fused(small)                            # plan complete, batched re-pack active
    mb.eval(); mf.eval()
    with torch.no_grad():
        lb, lf = mb(big["input"]), mf(big["input"])
    assert torch.isfinite(lb).all()
    assert (torch.sigmoid(lb) - torch.sigmoid(lf)).abs().max().item() <= 5e-2
    mb.train(); mf.train()
    l_b, l_f = mb.train_step_fused(big), mf.train_step_fused(big)
    assert abs(l_b.item() - l_f.item()) <= 3e-2 * abs(l_f.item())
    assert torch.isfinite(mb.network.flat_params).all() and torch.isfinite(mb.network.flat_grads).all()
    gb = dict(mb.network.named_parameters())
    gf = dict(mf.network.named_parameters())
    cos = lambda a, b: torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()
    i = list(gb).index("segmentation_head.0.weight")
    assert cos(mb.network._grad_views[i], mf.network._grad_views[i]) > 0.99


def test_normalizer_cache_follows_load_state_dict():
    """the fused normalise kernel must see parameters loaded / edited in place."""
    _, m = build_pair(1.0)
    m.eval()
    b = to_dev(synthetic.hyperstarcop_batch(1, size=64, seed=26))
    with torch.no_grad():
        a = m.normalizer.normalize_x(b["input"]).clone()
        sd = m.state_dict()
        sd["normalizer.factors_input"] = sd["normalizer.factors_input"] * 2
        m.load_state_dict(sd)
        c = m.normalizer.normalize_x(b["input"])
    ref = torch.clamp(b["input"] / (torch.tensor([1750., 60., 60., 60.], device=DEV) * 2)[None, :, None, None], 0, 2)
    assert not torch.equal(a, c)
    assert torch.allclose(c, ref, rtol=1e-6, atol=0)


def test_graphed_step_prefetch_from_pinned_host_matches_direct_stepping():
    """double-buffered inputs: the H2D copy of batch i+1 overlaps step i; every step must see ITS batch.
    Batch i carries weight_loss scaled by (i+1)^2, so its loss is ~(i+1)^2 x the base loss: a stale or swapped
    input buffer shows up as a factor, far above the ~0.5 % run-to-run drift of two bf16 training runs
    (fp32-atomic summation order in the weight gradients, BatchNorm over 8 pixels at this tile size)."""
    torch.manual_seed(7)
    m1 = get_model(default_settings(pos_weight=1.0, compute_dtype="bf16"), None).to(DEV)
    torch.manual_seed(7)
    m2 = get_model(default_settings(pos_weight=1.0, compute_dtype="bf16"), None).to(DEV)
    m1.train(); m2.train()
    host = []
    for i in range(4):
        b = synthetic.hyperstarcop_batch(2, size=64, seed=60)           # same tiles: only the weights differ
        b["weight_loss"] = b["weight_loss"] * float((i + 1) ** 2)
        host.append(b)
    pinned = [{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in b.items()} for b in host]
    s1 = m1.make_graphed_train_step(to_dev(host[0]), warmup=1)
    s2 = m2.make_graphed_train_step(to_dev(host[0]), warmup=1, double_buffer=True)
    l1 = [s1(to_dev(b)).item() for b in host]
    l2 = []
    s2.prefetch(pinned[0])
    for i in range(4):
        if i + 1 < 4:
            s2.prefetch(pinned[i + 1])
        l2.append(s2().item())
    for i, (a, b) in enumerate(zip(l1, l2)):
        assert abs(a - b) <= 2e-2 * abs(a), (i, a, b)
    assert all(l2[i + 1] > 1.5 * l2[i] for i in range(3)), l2      # ((i+2)/(i+1))^2 >= 16/9
    with pytest.raises(AssertionError):
        s2()                                   # nothing prefetched


def test_per_tile_validation_and_padded_predict():
    from oracle import loss_metrics as olm
    from starcop_b200 import tiling, validation
    oracle, model = build_pair(1.0)
    oracle.eval(); model.eval()
    tiles = [synthetic.hyperstarcop_batch(1, size=64, seed=40 + i) for i in range(3)]
    rows, gcm, sweep = validation.run_validation(model, tiles)
    assert len(rows) == 3 and len(sweep) == 16
    tot = torch.zeros(2, 2, dtype=torch.long)
    for t, row in zip(tiles, rows):
        with torch.no_grad():
            bo = oracle.batch_with_preds(t)
        safe = bo["logits"].abs() > 1e-3
        cm_o = olm.confusion_matrix(bo["pred_binary"], bo["output_norm"].long())
        # pixels whose oracle logit lies within the fp32 tolerance of 0 may flip: they bound the count difference
        unsafe = int((~safe).sum())
        assert abs(row["TP"] - int(cm_o[1, 1])) <= unsafe and abs(row["FP"] - int(cm_o[0, 1])) <= unsafe
        assert abs(row["FN"] - int(cm_o[1, 0])) <= unsafe and abs(row["TN"] - int(cm_o[0, 0])) <= unsafe
        assert row["TP"] + row["FP"] + row["FN"] + row["TN"] == 64 * 64
        assert row["label_pixels_plume"] == int(t["output"].sum())
        assert row["pred_classification"] == int(bo["pred_classification"][0, 0])
        tot += cm_o
    assert int(gcm.sum()) == 3 * 64 * 64
    # whole-scene prediction with reflect padding to a multiple of 32 (padding.py semantics)
    scene = synthetic.hyperstarcop_batch(1, size=96, seed=7)["input"][0][:, :70, :90]
    pred = tiling.padded_predict(scene.numpy(), model)
    assert pred.shape == (1, 70, 90)
    from oracle.tiling import find_padding
    pr, pc = find_padding(70, 32), find_padding(90, 32)
    padded = torch.nn.functional.pad(scene[None], (pc[0], pc[1], pr[0], pr[1]), mode="reflect")
    with torch.no_grad():
        ref = oracle(padded)[0][:, pr[0]:pr[0] + 70, pc[0]:pc[0] + 90]
    assert np.abs(pred - ref.numpy()).max() <= 1e-3
