"""Restatement of ``starcop/data/aviris.py:262-338`` ``transform_to_srf`` (without the georeader containers and the
final resize): pandas / scipy as the reference writes it.  Test infrastructure only (see ``oracle/__init__.py``)."""
import numpy as np
import pandas as pd
from scipy import interpolate


def transform_to_srf(cube_bsq, bands, srf, bands_nanometers, fill_value_default=0.0):
    """cube_bsq: (C, H, W) float32; srf: DataFrame indexed by wavelength with one column per band name."""
    bands_index = np.arange(0, len(bands_nanometers))
    interp = interpolate.interp1d(bands_nanometers, bands_index, kind="nearest")            # :277
    y_nearest = interp(srf.index).astype(int)
    table = pd.DataFrame({"SR_WL": srf.index, "AVIRIS_band": y_nearest}).set_index("SR_WL")
    out = np.full((len(bands),) + cube_bsq.shape[-2:], fill_value=fill_value_default, dtype=np.float32)
    for i, column_name in enumerate(bands):
        mask_zero = srf[column_name] <= 1e-4                                                 # :295
        w = srf.loc[~mask_zero, [column_name]].copy().join(table)
        norm = f"{column_name}_norm"
        w[norm] = w[column_name] / w[column_name].sum()                                      # :307-308
        per_band = w.groupby("AVIRIS_band")[[norm]].sum()
        sel = cube_bsq[per_band.index.values]
        missing = np.any(sel == fill_value_default, axis=0)                                  # :320-322
        out[i] = np.sum(per_band[norm].values[:, np.newaxis, np.newaxis] * sel, axis=0)      # :324-326
        out[i][missing] = fill_value_default
    return out
