"""Per-tile evaluation block of ``starcop/validation.py:80-135`` (run_validation): for each tile the
confusion matrix, the metric dictionary, pixel counts, tile classification and the 16-threshold
precision/recall sweep -- from the fused per-pixel products of ``ModelModule.batch_with_preds``.
Plotting, CSV/JSON dumps and the pandas aggregation of the reference are out of scope."""
import numpy as np
import torch

from . import metrics

DEFAULT_THRESHOLDS = [0, 1e-3, 1e-2] + np.arange(0.5, .96, .05).tolist() + [.99, .995, .999]   # validation.py:37


def _cm(pred, target):
    idx = 2 * target.long().flatten() + pred.long().flatten()
    return torch.bincount(idx, minlength=4).reshape(2, 2)


@torch.no_grad()
def run_validation(model, batches, thresholds=None):
    """batches: iterable of batch dicts with batch size 1 (validation.py:34).  Returns
    (per_tile: list of dicts, global_cm: 2x2 int64, sweep: list of (threshold, 2x2 int64))."""
    thresholds = np.sort(DEFAULT_THRESHOLDS if thresholds is None else thresholds)[::-1]
    model.eval()
    dev = next(model.parameters()).device
    sweep = [torch.zeros(2, 2, dtype=torch.long, device=dev) for _ in thresholds]
    global_cm = torch.zeros(2, 2, dtype=torch.long, device=dev)
    out = []
    for batch in batches:
        assert batch["input"].shape[0] == 1, "This function is expected to run with batch_size 1"
        b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
        b = model.batch_with_preds(b)
        y_long = b["output_norm"].long()
        cm = _cm(b["pred_binary"], y_long)
        global_cm += cm
        cmc = cm.cpu()
        row = {f.__name__: f(cmc).item() for f in metrics.METRICS_CONFUSION_MATRIX + [metrics.TP, metrics.TN, metrics.FP, metrics.FN]}
        for i, thr in enumerate(thresholds):                        # validation.py:118-125
            if hasattr(model, "apply_threshold"):
                pb = model.apply_threshold(b["prediction"], thr)
            else:
                pb = (b["prediction"] > thr).long()
            sweep[i] += _cm(pb, y_long)
        row["id"] = b["id"][0]
        row["label_pixels_plume"] = int(y_long[0, 0].sum().item())
        row["has_plume"] = int(torch.as_tensor(b["has_plume"]).reshape(-1)[0].item())
        row["pred_classification"] = int(b["pred_classification"][0, 0].item())
        row["pred_pixels_plume"] = int(b["pred_binary"][0, 0].sum().item())
        out.append(row)
    return out, global_cm.cpu(), [(float(t), s.cpu()) for t, s in zip(thresholds, sweep)]
