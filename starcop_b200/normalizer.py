"""DataNormalizer of the reference (starcop/data/normalizer_module.py) on the CUDA kernel
``sc_normalize_pack``: same constructor argument (the settings tree), same parameter names (so the
four ``normalizer.*_input`` entries of a reference state_dict load), same dtype quirks."""
import warnings

import numpy as np
import torch

from . import _lib


def _e(factor, clip=(0, 2), offset=0):
    return {"offset": offset, "factor": factor, "clip": clip}


# normalizer_module.py:7-74 (python int vs float literals are significant: they pick int64 / float64)
BAND_NORMALIZATION = {f"TOA_{s}_{b}": _e(1) for s in ("S2A", "S2B")
                      for b in ("B1", "B2", "B3", "B4", "B5", "B6", "B7", "B8", "B8A", "B9", "B10", "B11", "B12")}
BAND_NORMALIZATION.update({f"TOA_WV3_SWIR{i}": _e(1) for i in range(1, 9)})
BAND_NORMALIZATION.update({
    "TOA_AVIRIS_550nm": _e(60), "TOA_AVIRIS_640nm": _e(60), "TOA_AVIRIS_460nm": _e(60),
    "TOA_AVIRIS_2004nm": _e(1), "TOA_AVIRIS_2109nm": _e(5), "TOA_AVIRIS_2310nm": _e(4),
    "TOA_AVIRIS_2350nm": _e(3), "TOA_AVIRIS_2360nm": _e(3), "mag1c": _e(1750),
    "ratio_aviris_2350_2310_out": _e(0.0625, (-2., 2.)), "ratio_aviris_2350_2360_out": _e(0.0625, (-2., 2.)),
    "ratio_aviris_2360_2310_out": _e(0.0625, (-2., 2.)),
    "ratio_wv3_B7_B5_varon21_sum_c_out": _e(0.04, (-2., 2.)), "ratio_wv3_B8_B5_varon21_sum_c_out": _e(0.1, (-2., 2.)),
    "ratio_wv3_B7_B6_varon21_sum_c_out": _e(0.1, (-2., 2.)),
    "ratio_wv3_B7_B7MLR_SanchezGarcia22_sum_c_out": _e(0.025, (-2., 2.)),
    "ratio_wv3_B8_B8MLR_SanchezGarcia22_sum_c_out": _e(0.0769, (-2., 2.)),
    "ratio_wv3_B7_B7MLR_SanchezGarcia22_simplediv": _e(1, (-2., 2.)),
    "ratio_wv3_B8_B8MLR_SanchezGarcia22_simplediv": _e(1, (-2., 2.), -0.5),
    "ratio_lrn_bands2band8only_60ep_512_l1": _e(0.5, (-2., 2.)),
    "ratio_wv3_B7_B7MLR_fromS2_9bands_sum_c_out": _e(1, (-2., 2.)),
    "ratio_wv3_B7_B7MLR_fromS2_5bands_sum_c_out": _e(0.1111111, (-2., 2.)),
    "ratio_wv3_B8_B8MLR_fromS2_9bands_sum_c_out": _e(0.125, (-2., 2.)),
    "ratio_wv3_B8_B8MLR_fromS2_5bands_sum_c_out": _e(0.1666666, (-2., 2.)),
})


class DataNormalizer(torch.nn.Module):
    def __init__(self, settings):
        super().__init__()
        self.settings_dataset = settings.dataset
        off, fac, lo, hi = [], [], [], []
        for p in self.settings_dataset.input_products:
            if p not in BAND_NORMALIZATION:          # normalizer_module.py:87-93
                warnings.warn(f"Feature {p} does not have band normalization attributes. "
                              f"It will not be normalized BUT it will be clipped to [-10, 10]")
                off.append(0); fac.append(1); lo.append(-10); hi.append(10)
            else:
                e = BAND_NORMALIZATION[p]
                off.append(e["offset"]); fac.append(e["factor"]); lo.append(e["clip"][0]); hi.append(e["clip"][1])
        mk = lambda v: torch.nn.Parameter(torch.from_numpy(np.array(v)[:, None, None]), requires_grad=False)
        self.offsets_input, self.factors_input = mk(off), mk(fac)
        self.clip_min_input, self.clip_max_input = mk(lo), mk(hi)
        out = [p for p in self.settings_dataset.output_products if p in BAND_NORMALIZATION]
        if out:
            raise NotImplementedError("normalised output products are outside the HyperSTARCOP hot path "
                                      "(labelbinary is passed through, normalizer_module.py:140-144)")
        self.factors_output = None
        self.offsets_output = None
        self._dev_params = None

    def kernel_params(self, device):
        """(4,C) float64 device tensor [off, fac, lo, hi] + the float64 promotion mask."""
        # keyed on the parameters' storage AND version counters: load_state_dict / in-place edits invalidate it
        prm = (self.offsets_input, self.factors_input, self.clip_min_input, self.clip_max_input)
        key = (str(device),) + tuple((t.data_ptr(), t._version) for t in prm)
        if self._dev_params is None or self._dev_params[0] != key:
            arr = torch.stack([t.detach().reshape(-1).to(torch.float64) for t in
                               (self.offsets_input, self.factors_input, self.clip_min_input, self.clip_max_input)])
            mask = (1 if self.offsets_input.dtype == torch.float64 else 0) | \
                   (2 if self.factors_input.dtype == torch.float64 else 0)
            self._dev_params = (key, arr.to(device).contiguous(), mask)
        return self._dev_params[1], self._dev_params[2]

    def normalize_x(self, x):
        """clamp((x-offset)/factor, clip_min, clip_max).float() -- normalizer_module.py:134-135."""
        if not x.is_cuda:
            raise _lib.StarcopB200Error("starcop_b200 runs on CUDA tensors only (no CPU path)")
        squeeze = x.dim() == 3
        xx = (x[None] if squeeze else x).contiguous().float()
        B, C, H, W = xx.shape
        prm, mask = self.kernel_params(xx.device)
        assert prm.shape[1] == C, f"expected {prm.shape[1]} channels, got {C}"
        out = torch.empty_like(xx)
        st = torch.cuda.current_stream(xx.device).cuda_stream
        _lib.call("sc_normalize_pack", xx.data_ptr(), prm[0].data_ptr(), prm[1].data_ptr(), prm[2].data_ptr(),
                  prm[3].data_ptr(), mask, B, C, H, W, 0, C, _lib.SC_F32, out.data_ptr(), st)
        return out[0] if squeeze else out

    def denormalize_x(self, x):
        return (x * self.factors_input) + self.offsets_input

    def normalize_y(self, y):
        return y

    def denormalize_y(self, y):
        return y
