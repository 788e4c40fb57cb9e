"""Restatement of the loss / decision / confusion-matrix / metric arithmetic:

* weighted BCE-with-logits -- ``starcop/models/model_module.py:53-58,76-79``
  (``torch.nn.BCEWithLogitsLoss(pos_weight, reduction="none")`` then ``mean(loss*w)``)
* decision rules -- ``model_module.py:124`` (val: logits >= 0), ``:193,204`` (sigmoid > .5),
  ``:210-212`` pred_classification, ``:268-269`` differences
* torchmetrics 0.10 ``ConfusionMatrix(num_classes=2)`` update = bincount(2*target+pred)
* ``starcop/metrics.py:20-85``

Test infrastructure only (see ``oracle/__init__.py``).
"""
import numpy as np
import torch


def bce_with_logits_elementwise(logits, y, pos_weight):
    """ATen binary_cross_entropy_with_logits, reduction none:
    (1-y)*x + (1+(p-1)*y) * (log1p(exp(-|x|)) + max(-x,0))."""
    lw = 1 + (pos_weight - 1) * y
    return (1 - y) * logits + lw * (torch.log1p(torch.exp(-logits.abs())) + torch.clamp(-logits, min=0))


def weighted_bce(logits, y, w, pos_weight):
    """model_module.py:76-79."""
    return torch.mean(bce_with_logits_elementwise(logits, y, pos_weight) * w)


def weighted_bce_grad(logits, y, w, pos_weight):
    """d mean(l*w) / d logits in ATen's own form (binary_cross_entropy_with_logits_backward):
    ((p*y + 1 - y) * sigmoid(x) - p*y) * w / N -- same rounding in the tails as autograd."""
    t = pos_weight * y
    return ((t + 1 - y) * torch.sigmoid(logits) - t) * w / logits.numel()


def pred_val(logits):
    """model_module.py:124."""
    return (logits >= 0).long()


def pred_binary_sigmoid(logits):
    """model_module.py:193,204."""
    return (torch.sigmoid(logits) > .5).long()


def pred_classification(pred_binary):
    """model_module.py:210-212."""
    n_pixels = (10 * np.prod(tuple(pred_binary.shape[-2:]))) / (64 ** 2)
    return (torch.sum(pred_binary, dim=(-1, -2)) > n_pixels).long()


def differences(pred_binary, y_gt):
    """model_module.py:268-269 -- 0 TN, 1 FN, 2 FP, 3 TP."""
    return 2 * pred_binary.long() + (y_gt == 1).long()


def confusion_matrix(pred, target):
    """torchmetrics 0.10 binary confusion matrix: cm[target, pred], int64."""
    idx = (2 * target.long().flatten() + pred.long().flatten())
    return torch.bincount(idx, minlength=4).reshape(2, 2)


# ---- starcop/metrics.py -------------------------------------------------------------------
def precision(cm): return cm[1, 1] / (cm[1, 1] + cm[0, 1])            # :20-24
def recall(cm): return cm[1, 1] / (cm[1, 1] + cm[1, 0])               # :26-30
def f1score(cm):                                                      # :32-35
    p, r = precision(cm), recall(cm)
    return 2 * (p * r) / (p + r)
def FPR(cm): return cm[0, 1] / (cm[0, 1] + cm[0, 0])                  # :37-39
def iou(cm): return cm[1, 1] / (cm[1, 1] + cm[1, 0] + cm[0, 1])       # :41-45
def accuracy(cm): return (cm[1, 1] + cm[0, 0]) / cm.sum()             # :47-51
def cohen_kappa(cm):                                                  # :53-64
    c = cm.float() if not cm.is_floating_point() else cm
    s0, s1 = c.sum(dim=0, keepdim=True), c.sum(dim=1, keepdim=True)
    expected = s1 @ s0 / s0.sum()
    w = torch.ones_like(c).flatten(); w[::3] = 0; w = w.reshape(2, 2)
    return 1 - torch.sum(w * c) / torch.sum(w * expected)
def balanced_accuracy(cm):                                            # :66-71
    return 0.5 * (recall(cm) + cm[0, 0] / (cm[0, 0] + cm[0, 1]))
def TP(cm): return cm[1, 1]
def TN(cm): return cm[0, 0]
def FP(cm): return cm[0, 1]
def FN(cm): return cm[1, 0]

METRICS_CONFUSION_MATRIX = [precision, recall, f1score, iou, accuracy, cohen_kappa, balanced_accuracy]
