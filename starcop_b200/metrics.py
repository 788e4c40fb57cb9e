"""Confusion-matrix metrics with the names and semantics of ``starcop/metrics.py:8-85``.

``cm`` is the 2x2 int64 matrix ``cm[true, pred]`` = [[TN, FP], [FN, TP]] produced exactly (integer
counts) by the fused loss/metrics kernel; the ratios are formed on the host like the reference
does, including its behaviour on empty denominators (integer / integer true division -> NaN).
"""
import torch


def _chk(cm):
    assert cm.shape == (2, 2), f"Expected binary found {cm.shape}"


def TP(cm): return cm[1, 1]
def TN(cm): return cm[0, 0]
def FP(cm): return cm[0, 1]
def FN(cm): return cm[1, 0]


def precision(cm):
    _chk(cm)
    return TP(cm) / (TP(cm) + FP(cm))


def recall(cm):
    _chk(cm)
    return TP(cm) / (TP(cm) + FN(cm))


user_accuracy = precision
producer_accuracy = recall
TPR = recall


def f1score(cm):
    p, r = precision(cm), recall(cm)
    return 2 * (p * r) / (p + r)


def FPR(cm):
    return FP(cm) / (FP(cm) + TN(cm))


def iou(cm):
    _chk(cm)
    return TP(cm) / (TP(cm) + FN(cm) + FP(cm))


def accuracy(cm):
    _chk(cm)
    return (TP(cm) + TN(cm)) / cm.sum()


def cohen_kappa(cm):
    c = cm if cm.is_floating_point() else cm.float()
    col, row = c.sum(dim=0, keepdim=True), c.sum(dim=1, keepdim=True)
    expected = row @ col / col.sum()
    off_diag = 1 - torch.eye(2, dtype=c.dtype, device=c.device)
    return 1 - torch.sum(off_diag * c) / torch.sum(off_diag * expected)


def balanced_accuracy(cm):
    return 0.5 * (recall(cm) + TN(cm) / (TN(cm) + FP(cm)))


METRICS_CONFUSION_MATRIX = [precision, recall, f1score, iou, accuracy, cohen_kappa, balanced_accuracy]
