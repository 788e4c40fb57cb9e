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


@pytest.mark.parametrize("pw", [1.0, 15.0])
def test_train_step_fp32_matches_oracle_and_reference(golden, pw):
    g = golden("model_module.npz")
    oracle, model = build_pair(pw)
    batch = synthetic.hyperstarcop_batch(2, size=64, seed=3)
    oracle.train(); model.train()
    lo = oracle.training_step(batch, 0); lo.backward()
    lm = model.training_step(to_dev(batch), 0); lm.backward()
    tag = f"pw{int(pw)}"
    # loss: fp32 tolerance 1e-5 relative (reference golden and oracle)
    assert abs(lm.item() - float(g[f"{tag}_train_loss"])) <= 1e-5 * abs(float(g[f"{tag}_train_loss"]))
    assert abs(lm.item() - lo.item()) <= 1e-5 * abs(lo.item())
    # every parameter gradient, max-abs error relative to the gradient's max magnitude
    worst = 0.0
    for (n, po), (_, pm) in zip(oracle.network.named_parameters(), model.network.named_parameters()):
        assert pm.grad is not None, n
        e = rel_err(pm.grad.cpu(), po.grad)
        worst = max(worst, e)
        assert e <= 2e-3, (n, e)
    assert np.allclose(model.network.segmentation_head[0].weight.grad.cpu().numpy(), g[f"{tag}_grad_head_w"], rtol=1e-3, atol=1e-6)
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
        assert abs(lf.item() - lo.item()) <= 2e-4 * abs(lo.item()), (step, lf.item(), lo.item())


def test_bf16_mode_tracks_fp32_within_stated_tolerance():
    """bf16 storage / fp32 accumulate: logits within 5e-2 * max|logit| of the fp32 oracle after a
    training-mode forward (BN batch statistics), loss within 2e-2 relative."""
    oracle, model = build_pair(1.0, compute_dtype="bf16")
    batch = synthetic.hyperstarcop_batch(2, size=64, seed=3)
    oracle.train(); model.train()
    lo = oracle.training_step(batch, 0)
    lm = model.training_step(to_dev(batch), 0)
    lm.backward()
    assert abs(lm.item() - lo.item()) <= 2e-2 * abs(lo.item())
    lo.backward()
    go = oracle.network.decoder.blocks[0].conv1[0].weight.grad
    gm = model.network.decoder.blocks[0].conv1[0].weight.grad.cpu()
    cos = torch.nn.functional.cosine_similarity(go.flatten(), gm.flatten(), dim=0).item()
    assert cos > 0.98, cos


def test_requires_cuda_and_divisible_by_32():
    _, model = build_pair(1.0)
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 4, 48, 64, device=DEV))
    with pytest.raises(Exception):
        model(torch.zeros(1, 4, 64, 64))           # CPU tensor: no CPU path
