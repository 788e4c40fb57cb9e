"""Restatement of ``starcop/baselines.py``: ``Mag1cBaseline`` (:31-77), ``SanchezBaseline`` (:81-139),
``VaronBaseline`` (:142-200) ``batch_with_preds`` / ``apply_threshold`` on torch CPU, with the oracle's
``binary_opening`` (kornia restatement, oracle/morphology.py).  Test infrastructure only (see ``oracle/__init__.py``)."""
from . import loss_metrics as lm
from .morphology import binary_opening
from .normalizer import normalize_x, normalize_y


class ThresholdBaseline:
    def __init__(self, input_products, band_name, threshold, use_normalisation=True, use_morphological_ops=True):
        self.input_products = list(input_products)
        self.band = self.input_products.index(band_name)
        self.threshold = threshold
        self.use_normalisation = use_normalisation
        self.use_morphological_ops = use_morphological_ops

    def apply_threshold(self, pred, threshold):                      # :54-58 / :109-116 / :170-177
        t = pred > threshold
        return binary_opening(t).long() if self.use_morphological_ops else t.long()

    def batch_with_preds(self, batch):                               # :61-77 / :119-139 / :180-200
        batch = dict(batch)
        batch["input_norm"] = normalize_x(batch["input"], self.input_products)
        batch["output_norm"] = normalize_y(batch["output"], ["labelbinary"])
        src = batch["input_norm"] if self.use_normalisation else batch["input"]
        pred = src[:, self.band:self.band + 1]
        batch["prediction"] = pred
        batch["pred_binary"] = self.apply_threshold(pred, self.threshold)
        batch["differences"] = lm.differences(batch["pred_binary"], batch["output_norm"].long())
        batch["pred_classification"] = lm.pred_classification(batch["pred_binary"])
        return batch


def Mag1cBaseline(input_products, mag1c_threshold=500.0):
    return ThresholdBaseline(input_products, "mag1c", mag1c_threshold, use_normalisation=False)


def SanchezBaseline(input_products, baseline_threshold=0.05, use_normalisation=True, use_morphological_ops=True,
                    band_name="ratio_wv3_B8_B8MLR_SanchezGarcia22_sum_c_out"):
    return ThresholdBaseline(input_products, band_name, baseline_threshold, use_normalisation, use_morphological_ops)


def VaronBaseline(input_products, baseline_threshold=0.05, use_normalisation=True, use_morphological_ops=True):
    return ThresholdBaseline(input_products, "ratio_wv3_B7_B5_varon21_sum_c_out", baseline_threshold, use_normalisation,
                             use_morphological_ops)
