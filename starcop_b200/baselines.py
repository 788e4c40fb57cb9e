"""Drop-in for ``starcop/baselines.py``: the three threshold comparators of the paper -- ``Mag1cBaseline`` (:31-77),
``SanchezBaseline`` (:81-139), ``VaronBaseline`` (:142-200) -- with the reference's constructor arguments and method
surface (``forward``, ``apply_threshold``, ``batch_with_preds``, ``.normalizer``), so ``run_validation`` treats them
like a ``ModelModule`` (validation.py:80-135 calls ``batch_with_preds`` and, for the threshold sweep,
``apply_threshold``).  Threshold + opening with the 3x3 cross (``binary_opening``, :25-27, kornia erosion/dilation
with the default geodesic border) is ONE C-ABI call, ``sc_threshold_opening``; the tile counts and ``differences``
come from the fused reduction kernel's integer path.  CUDA tensors only (no CPU path)."""
from types import SimpleNamespace

import torch

from . import _lib
from .model_module import differences, pred_classification
from .normalizer import DataNormalizer


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def binary_opening_threshold(pred, threshold):
    """(pred > threshold) opened with the 3x3 cross -> int64 mask of pred's shape (baselines.py:25-27, 54-58)."""
    if not pred.is_cuda:
        raise _lib.StarcopB200Error("starcop_b200 runs on CUDA tensors only (no CPU path)")
    p = pred.contiguous().float()
    H, W = p.shape[-2:]
    B = p.numel() // (H * W)
    out = torch.empty(p.shape, dtype=torch.long, device=p.device)
    scratch = torch.empty(p.numel(), dtype=torch.uint8, device=p.device)
    _lib.call("sc_threshold_opening", p.data_ptr(), float(threshold), out.data_ptr(), scratch.data_ptr(), B, H, W, _stream(p.device))
    return out


class _ThresholdBaseline(torch.nn.Module):
    """shared body of the three reference classes (their batch_with_preds / apply_threshold are the same code)"""

    def __init__(self, input_products, band_name, threshold, use_normalisation, use_morphological_ops):
        super().__init__()
        self.band_baseline = list(input_products).index(band_name)
        self.baseline_threshold = threshold
        # the reference keeps the structuring element as a frozen parameter (it shows up in state_dict())
        self.element_stronger = torch.nn.Parameter(torch.tensor([[0., 1., 0.], [1., 1., 1.], [0., 1., 0.]]), requires_grad=False)
        settings_normalizer = SimpleNamespace(dataset=SimpleNamespace(input_products=list(input_products),
                                                                      output_products=["labelbinary"]))
        self.normalizer = DataNormalizer(settings_normalizer)
        self.use_normalisation = use_normalisation
        self.use_morphological_ops = use_morphological_ops

    @property
    def device(self):
        return self.element_stronger.device

    def forward(self, x):
        return x[:, self.band_baseline:(self.band_baseline + 1)]

    def apply_threshold(self, pred, threshold):
        if self.use_morphological_ops:
            return binary_opening_threshold(pred, threshold)
        return (pred > threshold).long()

    def batch_with_preds(self, batch):
        batch = batch.copy()
        batch["input_norm"] = self.normalizer.normalize_x(batch["input"])
        batch["output_norm"] = self.normalizer.normalize_y(batch["output"])
        pred = self(batch["input_norm"] if self.use_normalisation else batch["input"])
        batch["prediction"] = pred
        batch["pred_binary"] = self.apply_threshold(pred, self.baseline_threshold)
        batch["differences"] = differences(batch["pred_binary"], batch["output_norm"].long())
        batch["pred_classification"] = pred_classification(batch["pred_binary"])
        return batch


class Mag1cBaseline(_ThresholdBaseline):
    """baselines.py:31-77: threshold 500 ppm*m on the RAW mag1c band (no normalisation), always opened."""

    def __init__(self, input_products, mag1c_threshold=500.0):
        super().__init__(input_products, "mag1c", mag1c_threshold, use_normalisation=False, use_morphological_ops=True)
        self.band_mag1c = self.band_baseline
        self.mag1c_threshold = mag1c_threshold


class SanchezBaseline(_ThresholdBaseline):
    """baselines.py:81-139: B8 against the MLR of B1-B6 (Sanchez-Garcia 22), threshold 0.05 on the normalised band."""

    def __init__(self, input_products, baseline_threshold=0.05, use_normalisation=True, use_morphological_ops=True,
                 band_name="ratio_wv3_B8_B8MLR_SanchezGarcia22_sum_c_out"):
        super().__init__(input_products, band_name, baseline_threshold, use_normalisation, use_morphological_ops)


class VaronBaseline(_ThresholdBaseline):
    """baselines.py:142-200: the B7 / B5 ratio (Varon 21), threshold 0.05 on the normalised band."""

    def __init__(self, input_products, baseline_threshold=0.05, use_normalisation=True, use_morphological_ops=True):
        super().__init__(input_products, "ratio_wv3_B7_B5_varon21_sum_c_out", baseline_threshold, use_normalisation,
                         use_morphological_ops)
