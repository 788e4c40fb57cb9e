"""``starcop/validation.py`` ``run_validation`` without the plotting / file writing: the per-tile block (:80-135:
confusion matrix, metric dictionary, pixel counts, tile classification, 16-threshold sweep) from the fused
per-pixel products of ``batch_with_preds``, and the aggregation half (:155-222: difficulty split, global and
tile-classification metrics, precision / recall curve).  Works for ``ModelModule`` and for the threshold baselines
(``baselines.py``), whose ``apply_threshold`` (threshold + opening) drives the sweep like in the reference."""
import numpy as np
import torch

from . import _lib, metrics

DEFAULT_THRESHOLDS = [0, 1e-3, 1e-2] + np.arange(0.5, .96, .05).tolist() + [.99, .995, .999]   # validation.py:37


def _cm(pred, target):
    idx = 2 * target.long().flatten() + pred.long().flatten()
    return torch.bincount(idx, minlength=4).reshape(2, 2)


def _sweep_hist(prediction, y, thr_asc_dev, hist):
    """one pass over (prediction, y): histogram of 'number of thresholds exceeded' per class (sc_threshold_sweep)"""
    p, t = prediction.contiguous().float(), y.contiguous().float()
    _lib.call("sc_threshold_sweep", p.data_ptr(), t.data_ptr(), thr_asc_dev.data_ptr(), thr_asc_dev.numel(), p.numel(),
              hist.data_ptr(), torch.cuda.current_stream(p.device).cuda_stream)


def _sweep_cms(hist, K):
    """hist [2][K+1] -> list of K confusion matrices (ascending threshold order), exact integer prefix sums"""
    h = hist.cpu().reshape(2, K + 1)
    below = torch.cumsum(h, dim=1)[:, :K]              # pixels NOT above threshold k: exceeded <= k thresholds
    tot = h.sum(dim=1, keepdim=True)
    return [torch.stack([torch.stack([below[0, k], tot[0, 0] - below[0, k]]),
                         torch.stack([below[1, k], tot[1, 0] - below[1, k]])]) for k in range(K)]


@torch.no_grad()
def run_validation(model, batches, thresholds=None):
    """batches: iterable of batch dicts with batch size 1 (validation.py:34).  Returns
    (per_tile: list of dicts, global_cm: 2x2 int64, sweep: list of (threshold, 2x2 int64), thresholds high -> low)."""
    thresholds = np.sort(np.asarray(DEFAULT_THRESHOLDS if thresholds is None else thresholds, dtype=np.float64))[::-1]
    model.eval()
    dev = next(model.parameters()).device
    K = len(thresholds)
    fused = not hasattr(model, "apply_threshold")           # plain `prediction > thr`: all thresholds in one pass
    thr_asc = torch.as_tensor(np.ascontiguousarray(thresholds[::-1]).astype(np.float32), device=dev)
    hist = torch.zeros(2 * (K + 1), dtype=torch.long, device=dev)
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
        if fused:
            _sweep_hist(b["prediction"], b["output_norm"], thr_asc, hist)
        else:
            for i, thr in enumerate(thresholds):                        # validation.py:118-125
                sweep[i] += _cm(model.apply_threshold(b["prediction"], thr), y_long)
        row["id"] = b["id"][0]
        row["label_pixels_plume"] = int(y_long[0, 0].sum().item())
        row["has_plume"] = int(torch.as_tensor(b["has_plume"]).reshape(-1)[0].item()) if "has_plume" in b else int(row["label_pixels_plume"] > 0)
        row["pred_classification"] = int(b["pred_classification"][0, 0].item())
        row["pred_pixels_plume"] = int(b["pred_binary"][0, 0].sum().item())
        out.append(row)
    if fused:
        cms = _sweep_cms(hist, K)[::-1]                                  # back to high -> low
        return out, global_cm.cpu(), [(float(t), c) for t, c in zip(thresholds, cms)]
    return out, global_cm.cpu(), [(float(t), s.cpu()) for t, s in zip(thresholds, sweep)]


def aggregate(per_tile, global_cm, sweep):
    """validation.py:155-222 on plain Python containers (the reference uses a pandas groupby): per-tile rows gain
    ``has_plume`` (label pixels > 0) and ``difficulty`` ("easy" if label pixels > 1000 else "hard"); returns
    (rows, metrics) with FPR on plume-free tiles, the metric set per difficulty, the global pixel metrics, the
    tile-classification metrics and the thresholded precision / recall / TPR / FPR curve.  0/0 ratios stay NaN like
    the reference's tensor divisions; a missing (has_plume, difficulty) group contributes zeros."""
    rows = []
    groups = {}
    for r in per_tile:
        r = dict(r)
        r["has_plume"] = r["label_pixels_plume"] > 0
        r["difficulty"] = "easy" if r["label_pixels_plume"] > 1000 else "hard"
        g = groups.setdefault((r["has_plume"], r["difficulty"]), {"TP": 0, "FP": 0, "TN": 0, "FN": 0})
        for k in g:
            g[k] += int(r[k])
        rows.append(r)
    for g in groups.values():
        g["total"] = g["TP"] + g["FP"] + g["TN"] + g["FN"]
    grand = sum(g["total"] for g in groups.values())
    zero = {"TP": 0, "FP": 0, "TN": 0, "FN": 0, "total": 0}
    out = {}
    item = groups.get((False, "hard"), zero)
    out["FPR_no_plume"] = item["FP"] / (item["FP"] + item["TN"]) if (item["FP"] + item["TN"]) else float("nan")
    out["frac_total_easy"] = item["total"] / grand if grand else float("nan")     # (sic) validation.py:171 stores it under this key
    for d in ("easy", "hard"):
        item = groups.get((True, d), zero)
        # the reference reads these counts from a pandas row that also holds the float column frac_total: the
        # confusion matrix is a float64 tensor there, and so are the per-difficulty ratios
        cm_d = torch.tensor([[item["TN"], item["FP"]], [item["FN"], item["TP"]]], dtype=torch.float64)
        for f in metrics.METRICS_CONFUSION_MATRIX:
            out[f"{f.__name__}_{d}"] = f(cm_d).item()
        out[f"frac_total_{d}"] = item["total"] / grand if grand else float("nan")
    cm = torch.as_tensor(global_cm)
    for f in metrics.METRICS_CONFUSION_MATRIX:
        out[f.__name__] = f(cm).item()
    out["confusion_matrix"] = cm
    pc = torch.tensor([int(r["pred_classification"]) for r in rows], dtype=torch.long)
    hp = torch.tensor([int(r["has_plume"]) for r in rows], dtype=torch.long)
    cm_cls = torch.bincount(2 * hp + pc, minlength=4).reshape(2, 2) if len(rows) else torch.zeros(2, 2, dtype=torch.long)
    for f in metrics.METRICS_CONFUSION_MATRIX:
        out[f"classification_{f.__name__}"] = f(cm_cls).item()
    out["classification_confusion_matrix"] = cm_cls
    out["thresholded"] = []
    for thr, c in sweep:                                                   # thresholds from high to low
        d = {"threshold": thr, "confusion_matrix": c}
        for f in (metrics.precision, metrics.recall, metrics.TPR, metrics.FPR):
            d[f.__name__] = f(c)
        out["thresholded"].append(d)
    return rows, out
