"""Restatement of the aggregation half of ``starcop/validation.py`` ``run_validation`` (:155-222) with pandas, as
the reference writes it.  Test infrastructure only (see ``oracle/__init__.py``)."""
import pandas as pd
import torch

from . import loss_metrics as lm


def aggregate(per_tile, global_cm, sweep):
    out_data = pd.DataFrame(per_tile).set_index("id")
    out_data["has_plume"] = out_data["label_pixels_plume"] > 0                                    # :158
    out_data["difficulty"] = out_data["label_pixels_plume"].apply(lambda x: "easy" if x > 1000 else "hard")
    g = out_data.groupby(["has_plume", "difficulty"])[["TP", "FP", "TN", "FN"]].sum()
    g["total"] = g.sum(axis=1)
    g["frac_total"] = g["total"] / g["total"].sum()
    metrics = {}
    item = g.loc[(False, "hard")]
    metrics["FPR_no_plume"] = item.FP / (item.FP + item.TN)
    metrics["frac_total_easy"] = item.frac_total
    for d in ["easy", "hard"]:
        item = g.loc[(True, d)]
        cm_d = torch.tensor([[item.TN, item.FP], [item.FN, item.TP]], requires_grad=False)
        for f in lm.METRICS_CONFUSION_MATRIX:
            metrics[f"{f.__name__}_{d}"] = f(cm_d).item()
        metrics[f"frac_total_{d}"] = item.frac_total
    cm = torch.as_tensor(global_cm)
    for f in lm.METRICS_CONFUSION_MATRIX:
        metrics[f.__name__] = f(cm).item()
    pc = torch.from_numpy(out_data["pred_classification"].values).long()
    hp = torch.from_numpy(out_data["has_plume"].values).long()
    cm_cls = lm.confusion_matrix(pc, hp)
    for f in lm.METRICS_CONFUSION_MATRIX:
        metrics[f"classification_{f.__name__}"] = f(cm_cls).item()
    metrics["classification_confusion_matrix"] = cm_cls
    return out_data, metrics
