"""Restatement of ``starcop/data/normalizer_module.py`` (DataNormalizer).

Test infrastructure only (see ``oracle/__init__.py``).
"""
import numpy as np
import torch


def _bn(factor, clip=(0, 2), offset=0):
    return {"offset": offset, "factor": factor, "clip": clip}


# normalizer_module.py:7-74 -- same products, same python literal types (int vs float
# matters: np.array([...]) of python ints is int64, any float makes the array float64).
BAND_NORMALIZATION = {}
for _s in ("S2A", "S2B"):
    for _b in ("B1", "B2", "B3", "B4", "B5", "B6", "B7", "B8", "B8A", "B9", "B10", "B11", "B12"):
        BAND_NORMALIZATION[f"TOA_{_s}_{_b}"] = _bn(1)
for _i in range(1, 9):
    BAND_NORMALIZATION[f"TOA_WV3_SWIR{_i}"] = _bn(1)
BAND_NORMALIZATION.update({
    "TOA_AVIRIS_550nm": _bn(60), "TOA_AVIRIS_640nm": _bn(60), "TOA_AVIRIS_460nm": _bn(60),
    "TOA_AVIRIS_2004nm": _bn(1), "TOA_AVIRIS_2109nm": _bn(5), "TOA_AVIRIS_2310nm": _bn(4),
    "TOA_AVIRIS_2350nm": _bn(3), "TOA_AVIRIS_2360nm": _bn(3),
    "mag1c": _bn(1750),
    "ratio_aviris_2350_2310_out": _bn(0.0625, (-2., 2.)),
    "ratio_aviris_2350_2360_out": _bn(0.0625, (-2., 2.)),
    "ratio_aviris_2360_2310_out": _bn(0.0625, (-2., 2.)),
    "ratio_wv3_B7_B5_varon21_sum_c_out": _bn(0.04, (-2., 2.)),
    "ratio_wv3_B8_B5_varon21_sum_c_out": _bn(0.1, (-2., 2.)),
    "ratio_wv3_B7_B6_varon21_sum_c_out": _bn(0.1, (-2., 2.)),
    "ratio_wv3_B7_B7MLR_SanchezGarcia22_sum_c_out": _bn(0.025, (-2., 2.)),
    "ratio_wv3_B8_B8MLR_SanchezGarcia22_sum_c_out": _bn(0.0769, (-2., 2.)),
    "ratio_wv3_B7_B7MLR_SanchezGarcia22_simplediv": _bn(1, (-2., 2.)),
    "ratio_wv3_B8_B8MLR_SanchezGarcia22_simplediv": _bn(1, (-2., 2.), -0.5),
    "ratio_lrn_bands2band8only_60ep_512_l1": _bn(0.5, (-2., 2.)),
    "ratio_wv3_B7_B7MLR_fromS2_9bands_sum_c_out": _bn(1, (-2., 2.)),
    "ratio_wv3_B7_B7MLR_fromS2_5bands_sum_c_out": _bn(0.1111111, (-2., 2.)),
    "ratio_wv3_B8_B8MLR_fromS2_9bands_sum_c_out": _bn(0.125, (-2., 2.)),
    "ratio_wv3_B8_B8MLR_fromS2_5bands_sum_c_out": _bn(0.1666666, (-2., 2.)),
})


def normalizer_params(products):
    """normalizer_module.py:81-107 -- (offsets, factors, clip_min, clip_max), each (C,1,1)."""
    off, fac, lo, hi = [], [], [], []
    for p in products:
        if p not in BAND_NORMALIZATION:                  # :87-93
            off.append(0); fac.append(1); lo.append(-10); hi.append(10)
        else:
            e = BAND_NORMALIZATION[p]
            off.append(e["offset"]); fac.append(e["factor"]); lo.append(e["clip"][0]); hi.append(e["clip"][1])
    return tuple(torch.from_numpy(np.array(v)[:, None, None]) for v in (off, fac, lo, hi))


def normalize_x(x, products):
    """normalizer_module.py:134-135 -- clamp((x-off)/fac, lo, hi).float()."""
    off, fac, lo, hi = normalizer_params(products)
    return torch.clamp((x - off) / fac, lo, hi).float()


def normalize_y(y, output_products=("labelbinary",)):
    """normalizer_module.py:140-144 -- identity unless an output product is in the table."""
    if any(p in BAND_NORMALIZATION for p in output_products):
        off, fac, lo, hi = normalizer_params(output_products)
        return torch.clamp((y - off) / fac, lo, hi)
    return y
