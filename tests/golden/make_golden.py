"""Generate tests/golden/*.npz by running the REFERENCE's own code (read-only at
/root/reference) on seeded inputs.  Only runs in the build container; the vectors it
writes are committed and travel to the GPU box.

    python tests/golden/make_golden.py

Modules imported unmodified: starcop.metrics, starcop.data.normalizer_module,
starcop.data.feature_extration, starcop.models.mag1c (stub rasterio; ``spectral`` stub that
only implements the ENVI LUT read), starcop.models.model_module (stub pytorch_lightning,
torchmetrics = restated binary ConfusionMatrix, segmentation_models_pytorch.Unet = oracle.unet.Unet,
starcop.utils stubbed because it imports rasterio/fsspec at module scope).
"""
import os
import re
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle.unet import Unet                                   # noqa: E402
from starcop_b200 import synthetic                             # noqa: E402
from starcop_b200.settings import default_settings             # noqa: E402


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("rasterio"); mod("rasterio.windows")
    # ---- spectral: only spectral.io.envi.open(hdr, lut) -> .asarray(), .bands.centers
    def envi_open(hdr, lut):
        txt = open(hdr).read()
        wl = np.array([float(v) for v in re.search(r"wavelength = \{([^}]*)\}", txt, re.S).group(1).split(",")])
        samples = int(re.search(r"samples\s*=\s*(\d+)", txt).group(1))
        bands = int(re.search(r"bands\s*=\s*(\d+)", txt).group(1))
        data = np.fromfile(lut, "<f8").reshape(bands, 1, samples)          # BSQ: (bands, lines, samples)
        arr = np.transpose(data, (1, 2, 0))                                 # spectral: (lines, samples, bands)
        return types.SimpleNamespace(asarray=lambda: arr, bands=types.SimpleNamespace(centers=list(wl)))
    io = mod("spectral.io", envi=types.SimpleNamespace(open=envi_open))
    mod("spectral", io=io)
    mod("spectral.io.envi", open=envi_open)

    # ---- pytorch_lightning
    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k): pass
        def log(self, *a, **k): pass
    mod("pytorch_lightning", LightningModule=LightningModule)

    # ---- torchmetrics 0.10 binary ConfusionMatrix
    class ConfusionMatrix(torch.nn.Module):
        def __init__(self, num_classes=2, task=None):
            super().__init__()
            self.confmat = torch.zeros(2, 2, dtype=torch.long)
        def update(self, preds, target):
            idx = 2 * target.long().flatten() + preds.long().flatten()
            self.confmat += torch.bincount(idx, minlength=4).reshape(2, 2)
        def compute(self): return self.confmat.clone()
        def reset(self): self.confmat.zero_()
        def forward(self, preds, target):
            idx = 2 * target.long().flatten() + preds.long().flatten()
            cm = torch.bincount(idx, minlength=4).reshape(2, 2)
            self.confmat += cm
            return cm
    tmf = mod("torchmetrics.functional", mean_squared_error=lambda a, b: torch.mean((a - b) ** 2))
    mod("torchmetrics", ConfusionMatrix=ConfusionMatrix, functional=tmf)
    mod("segmentation_models_pytorch", Unet=Unet)
    # starcop.utils imports rasterio/fsspec/requests/pandas at module scope; only get_filesystem is used
    import starcop                                                          # noqa
    mod("starcop.utils", get_filesystem=lambda p: None)


def main():
    install_stubs()
    from starcop import metrics as ref_metrics
    from starcop.data import normalizer_module as ref_norm
    from starcop.data import feature_extration as ref_feat
    from starcop.models import mag1c as ref_mag1c
    from starcop.models import model_module as ref_mm

    out = lambda n: os.path.join(HERE, n)

    # ------------------------------------------------------------------ A1 normalizer
    g = torch.Generator().manual_seed(0)
    res = {}
    for tag, prods in (("hyper", ["mag1c", "TOA_AVIRIS_640nm", "TOA_AVIRIS_550nm", "TOA_AVIRIS_460nm"]),
                       ("multi", ["ratio_wv3_B7_B5_varon21_sum_c_out", "TOA_WV3_SWIR1",
                                  "ratio_wv3_B8_B8MLR_SanchezGarcia22_simplediv", "unknown_product"])):
        s = default_settings(input_products=prods)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            dn = ref_norm.DataNormalizer(s)
        scale = torch.tensor([4000., 150., 150., 150.])[:, None, None] if tag == "hyper" else torch.tensor([.3, 3., 3., 30.])[:, None, None]
        x = (torch.rand(2, 4, 16, 16, generator=g) - 0.1) * scale
        res[f"{tag}_x"] = x.numpy(); res[f"{tag}_y"] = dn.normalize_x(x).numpy()
        res[f"{tag}_factor_dtype"] = np.array(str(dn.factors_input.dtype))
    np.savez_compressed(out("normalizer.npz"), **res)

    # ------------------------------------------------------------------ A5 metrics
    cms = np.array([[[100, 5], [7, 30]], [[4096, 0], [0, 0]], [[0, 0], [0, 50]], [[10, 3], [0, 0]],
                    [[262000, 44], [60, 40]], [[1, 1], [1, 1]]], dtype=np.int64)
    names = [f.__name__ for f in ref_metrics.METRICS_CONFUSION_MATRIX] + ["TP", "TN", "FP", "FN", "FPR", "TPR"]
    vals = np.array([[float(getattr(ref_metrics, n)(torch.from_numpy(cm))) for n in names] for cm in cms])
    np.savez_compressed(out("metrics.npz"), cms=cms, names=np.array(names), values=vals)

    # ------------------------------------------------------------------ A3/A4/A6 model module
    res = {}
    for pw in (1.0, 15.0):
        torch.manual_seed(1234)
        s = default_settings(pos_weight=pw)
        mm = ref_mm.ModelModule(s)
        sd = {k: v.clone() for k, v in mm.state_dict().items()}
        batch = synthetic.hyperstarcop_batch(2, size=64, seed=3)
        mm.train()
        loss = mm.training_step(batch, 0)
        loss.backward()
        tag = f"pw{int(pw)}"
        res[f"{tag}_train_loss"] = loss.detach().numpy()
        res[f"{tag}_grad_head_w"] = mm.network.segmentation_head[0].weight.grad.numpy()
        res[f"{tag}_grad_stem_w"] = mm.network.encoder.features[0][0].weight.grad.numpy()
        res[f"{tag}_grad_dec_b4c2_bn_w"] = mm.network.decoder.blocks[4].conv2[1].weight.grad.numpy()
        mm.eval()
        with torch.no_grad():
            mm.val_step(batch, 0)
            res[f"{tag}_val_cm"] = mm.confusion_matrix.compute().numpy()
            res[f"{tag}_val_cm_cls"] = mm.classification_confusion_matrix.compute().numpy()
            bp = mm.batch_with_preds(batch)
        for k in ("logits", "prediction", "pred_binary", "differences", "pred_classification",
                  "loss_per_pixel", "loss_per_pixel_weighted", "input_norm"):
            res[f"{tag}_{k}"] = bp[k].numpy()
        if pw == 1.0:
            # weights are reproducible from the seed (the only RNG consumer in __init__ is the
            # network); pin them with per-tensor sums instead of committing 26 MB
            res["state_dict_sums"] = np.array([float(v.double().sum()) for v in sd.values()])
            res["state_dict_keys"] = np.array(list(sd.keys()))
            res["n_params"] = np.array(sum(p.numel() for p in mm.parameters()))
            res["n_trainable"] = np.array(sum(p.numel() for p in mm.parameters() if p.requires_grad))
    # elementwise loss + grad on adversarial logits (reference = torch BCEWithLogitsLoss itself)
    lg = torch.tensor([-100., -20., -1e-8, 0., 1e-8, 5e-8, 1.2e-7, 0.3, 20., 100.]).repeat(2)
    yy = torch.cat([torch.zeros(10), torch.ones(10)]); ww = torch.linspace(0.1, 1, 20)
    for pw in (1.0, 15.0):
        lgr = lg.clone().requires_grad_(True)
        fn = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor(pw), reduction="none")
        l = fn(lgr, yy); torch.mean(l * ww).backward()
        res[f"edge_pw{int(pw)}_loss"] = l.detach().numpy(); res[f"edge_pw{int(pw)}_grad"] = lgr.grad.numpy()
    res["edge_logits"], res["edge_y"], res["edge_w"] = lg.numpy(), yy.numpy(), ww.numpy()
    res["edge_pred_sigmoid"] = (torch.sigmoid(lg) > .5).long().numpy()
    res["edge_pred_val"] = (lg >= 0).long().numpy()
    pc_in = torch.zeros(3, 1, 64, 64, dtype=torch.long); pc_in[1, 0, :10, 0] = 1; pc_in[2, 0, :11, 0] = 1
    res["pred_classification_in"] = pc_in.numpy(); res["pred_classification_out"] = ref_mm.pred_classification(pc_in).numpy()
    np.savez_compressed(out("model_module.npz"), **res)

    # ------------------------------------------------------------------ A9 template
    wl = np.array(ref_feat.AVIRIS_WAVELENGTHS, dtype=np.float64)
    keep = ref_mag1c.get_mask_bad_bands(wl) & (wl > 2122) & (wl < 2488)      # process_aviris.py:192-195
    idx = np.where(keep)[0]
    centers = wl[keep]; fwhm = np.full_like(centers, 5.0)
    tmpl = ref_mag1c.generate_template_from_bands(centers, fwhm)
    np.savez_compressed(out("ch4_template_aviris.npz"), centers=centers, fwhm=fwhm, template=tmpl,
                        band_first=idx[0], band_last=idx[-1], keep_mask=ref_mag1c.get_mask_bad_bands(wl), wavelengths=wl)

    # ------------------------------------------------------------------ A8/A10 matched filter
    res = {}
    t73 = tmpl[:, 1]
    cube, _, alpha = synthetic.aviris_cube(1, size=128, bands=125, seed=5, template=t73)
    x = torch.from_numpy(np.ascontiguousarray(cube[0, :, 40:46, 52:125])).permute(1, 0, 2).contiguous()   # [b=6,p=128,s=73]
    res["x_f32"] = x.numpy(); res["template"] = t73
    for tag, dt, al in (("f32", torch.float32, 0.0), ("f64", torch.float64, 1e-4)):
        xx, tt = x.to(dt), torch.from_numpy(t73).to(dt)
        mf, R = ref_mag1c.rmf(xx, tt, alpha=al)
        res[f"rmf_{tag}_mf"], res[f"rmf_{tag}_R"] = mf.numpy(), R.numpy()
        for it in (1, 30):
            mf, R = ref_mag1c.acrwl1mf(xx, tt, num_iter=it, alpha=al)
            res[f"acrwl1mf_{tag}_it{it}_mf"], res[f"acrwl1mf_{tag}_it{it}_R"] = mf.numpy(), R.numpy()
    # func_by_groups with NODATA pixels and a <=10 px group
    cube2 = cube[0, :64, :12, 52:125].copy()
    groups = np.tile(np.arange(12)[None, :] // 2, (64, 1)).astype(np.int64)          # 6 groups of 2 columns
    cube2[3:9, 5, :] = ref_mag1c.NODATA
    mask = np.all(cube2 > ref_mag1c.NODATA, axis=-1)
    mask[:, 10:] = False; mask[:4, 10] = True                                        # group 5 has 4 valid px
    tt = torch.from_numpy(t73).float()
    mf, al_ = ref_mag1c.func_by_groups(lambda xg: ref_mag1c.acrwl1mf(xg, tt, num_iter=30), cube2, groups, mask, disable_pbar=True)
    res["fbg_cube"], res["fbg_groups"], res["fbg_mask"] = cube2, groups, mask
    res["fbg_mf"], res["fbg_albedo"] = mf.numpy(), al_.numpy()
    np.savez_compressed(out("mag1c.npz"), **res)

    # ------------------------------------------------------------------ A11-A13 features
    res = {}
    bg, sig = synthetic.ratio_bands(64, seed=2)
    res["bg"], res["sig"] = bg, sig
    res["ratio"] = ref_feat.ratio_2c_match_c_from_sums_outlier(bg.copy(), sig.copy())
    rng = np.random.default_rng(4)
    bands = [np.abs(bg * rng.uniform(0.5, 1.5) + rng.normal(0, 0.05, bg.shape)).astype(np.float32) for _ in range(5)]
    res["mlr_bands"] = np.stack(bands); res["mlr_target"] = sig
    res["mlr_ratio"] = ref_feat.ratio_MLR_local_5IN(*[b.copy() for b in bands], sig.copy())
    mag = rng.exponential(300, (32, 32)).astype(np.float32)
    res["mag1c"], res["weight_mag1c"] = mag, ref_feat.weight_mag1c(mag)
    np.savez_compressed(out("features.npz"), **res)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
