"""Pin the CPU oracle against vectors produced by the reference's own code
(tests/golden/make_golden.py, run in the build container against /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import features, loss_metrics as lm, mag1c, morphology, normalizer, tiling
from oracle.module import get_model
from oracle.unet import Unet, count_parameters
from starcop_b200 import synthetic
from starcop_b200.settings import default_settings


def test_unet_param_counts_match_reference_notebook():
    # notebooks/(bonus)_training_demo.ipynb:603-611: 6.6 M params, 26.517 MB incl. 17 frozen
    assert count_parameters(Unet(in_channels=4)) == 6629233
    assert count_parameters(Unet(in_channels=3)) == 6628945
    assert count_parameters(Unet(in_channels=1)) == 6628369
    assert round((6629233 + 17) * 4 / 1e6, 3) == 26.517


def test_unet_encoder_is_torchvision_mobilenet_v2():
    torchvision = pytest.importorskip("torchvision")
    from oracle.unet import MobileNetV2Encoder
    enc = MobileNetV2Encoder(3).eval()
    tv = torchvision.models.MobileNetV2().features.eval()
    tv.load_state_dict({k[len("features."):]: v for k, v in enc.state_dict().items()})
    x = torch.randn(1, 3, 64, 64)
    assert torch.equal(enc(x)[-1], tv(x))


def test_normalizer(golden):
    g = golden("normalizer.npz")
    prods = {"hyper": ["mag1c", "TOA_AVIRIS_640nm", "TOA_AVIRIS_550nm", "TOA_AVIRIS_460nm"],
             "multi": ["ratio_wv3_B7_B5_varon21_sum_c_out", "TOA_WV3_SWIR1",
                       "ratio_wv3_B8_B8MLR_SanchezGarcia22_simplediv", "unknown_product"]}
    for tag, p in prods.items():
        y = normalizer.normalize_x(torch.from_numpy(g[f"{tag}_x"]), p)
        assert y.dtype == torch.float32
        assert np.array_equal(y.numpy(), g[f"{tag}_y"])                    # bit exact
        assert str(normalizer.normalizer_params(p)[1].dtype) == str(g[f"{tag}_factor_dtype"])


def test_metrics(golden):
    g = golden("metrics.npz")
    for cm, vals in zip(g["cms"], g["values"]):
        for n, v in zip(g["names"], vals):
            fn = {"TPR": lm.recall}.get(str(n)) or getattr(lm, str(n))
            got = float(fn(torch.from_numpy(cm)))
            assert (np.isnan(got) and np.isnan(v)) or got == v, (cm, n, got, v)


@pytest.mark.parametrize("pw", [1, 15])
def test_model_module(golden, pw):
    g = golden("model_module.npz")
    torch.manual_seed(1234)
    mm = get_model(default_settings(pos_weight=float(pw)), None)
    if pw == 1:
        sums = np.array([float(v.double().sum()) for k, v in mm.network.state_dict().items()])
        ref = {str(k): s for k, s in zip(g["state_dict_keys"], g["state_dict_sums"])}
        for k, s in zip(mm.network.state_dict().keys(), sums):
            assert ref["network." + k] == s, k
        assert int(g["n_trainable"]) == count_parameters(mm.network) == 6629233
        assert int(g["n_params"]) == 6629233 + 17
    batch = synthetic.hyperstarcop_batch(2, size=64, seed=3)
    mm.train()
    loss = mm.training_step(batch, 0)
    loss.backward()
    t = f"pw{pw}"
    assert np.allclose(loss.item(), g[f"{t}_train_loss"], rtol=1e-6)
    assert np.allclose(mm.network.segmentation_head[0].weight.grad.numpy(), g[f"{t}_grad_head_w"], rtol=1e-4, atol=1e-7)
    assert np.allclose(mm.network.encoder.features[0][0].weight.grad.numpy(), g[f"{t}_grad_stem_w"], rtol=1e-3, atol=1e-6)
    mm.eval()
    with torch.no_grad():
        mm.val_step(batch, 0)
        assert np.array_equal(mm.cm.numpy(), g[f"{t}_val_cm"])
        assert np.array_equal(mm.cm_cls.numpy(), g[f"{t}_val_cm_cls"])
        bp = mm.batch_with_preds(batch)
    for k in ("pred_binary", "differences", "pred_classification"):
        assert np.array_equal(bp[k].numpy(), g[f"{t}_{k}"]), k
    for k in ("logits", "prediction", "loss_per_pixel", "loss_per_pixel_weighted", "input_norm"):
        assert np.allclose(bp[k].numpy(), g[f"{t}_{k}"], rtol=1e-5, atol=1e-6), k


def test_loss_edge_cases(golden):
    g = golden("model_module.npz")
    x, y, w = (torch.from_numpy(g[k]) for k in ("edge_logits", "edge_y", "edge_w"))
    for pw in (1, 15):
        l = lm.bce_with_logits_elementwise(x, y, torch.tensor(float(pw)))
        assert np.allclose(l.numpy(), g[f"edge_pw{pw}_loss"], rtol=1e-6, atol=1e-30)
        gr = lm.weighted_bce_grad(x, y, w, float(pw))
        assert np.allclose(gr.numpy(), g[f"edge_pw{pw}_grad"], rtol=1e-5, atol=1e-12)
    assert np.array_equal(lm.pred_binary_sigmoid(x).numpy(), g["edge_pred_sigmoid"])
    assert np.array_equal(lm.pred_val(x).numpy(), g["edge_pred_val"])
    # the two predicates differ for tiny positive logits
    assert (g["edge_pred_sigmoid"] != g["edge_pred_val"]).any()
    assert np.array_equal(lm.pred_classification(torch.from_numpy(g["pred_classification_in"])).numpy(),
                          g["pred_classification_out"])


def test_mag1c_band_selection(golden):
    g = golden("ch4_template_aviris.npz")
    assert np.array_equal(mag1c.get_mask_bad_bands(g["wavelengths"]), g["keep_mask"])
    sl = mag1c.band_keep_aviris(g["wavelengths"])
    assert (sl.start, sl.stop - 1) == (int(g["band_first"]), int(g["band_last"])) == (349, 421)
    assert sl.stop - sl.start == synthetic.AVIRIS_WINDOW_BANDS


@pytest.mark.parametrize("tag,dt,alpha,tol", [("f32", torch.float32, 0.0, 2e-3), ("f64", torch.float64, 1e-4, 1e-9)])
def test_mag1c_filters(golden, tag, dt, alpha, tol):
    g = golden("mag1c.npz")
    x = torch.from_numpy(g["x_f32"]).to(dt)
    t = torch.from_numpy(g["template"]).to(dt)
    mf, R = mag1c.rmf(x, t, alpha=alpha)
    assert np.allclose(R.numpy(), g[f"rmf_{tag}_R"], rtol=1e-6)
    scale = np.abs(g[f"rmf_{tag}_mf"]).max()
    assert np.abs(mf.numpy() - g[f"rmf_{tag}_mf"]).max() <= tol * scale
    for it in (1, 30):
        mf, R = mag1c.acrwl1mf(x, t, num_iter=it, alpha=alpha)
        ref = g[f"acrwl1mf_{tag}_it{it}_mf"]
        assert np.abs(mf.numpy() - ref).max() <= tol * np.abs(ref).max(), (it, np.abs(mf.numpy() - ref).max())
        assert np.allclose(R.numpy(), g[f"acrwl1mf_{tag}_it{it}_R"], rtol=1e-6)


def test_mag1c_func_by_groups(golden):
    g = golden("mag1c.npz")
    t = torch.from_numpy(g["template"]).float()
    mf, al = mag1c.func_by_groups(lambda xg: mag1c.acrwl1mf(xg, t, num_iter=30), g["fbg_cube"], g["fbg_groups"], g["fbg_mask"])
    ref = g["fbg_mf"]
    assert np.array_equal(mf.numpy() == mag1c.NODATA, ref == mag1c.NODATA)          # skipped / masked pixels bit exact
    assert (ref[:, 10:] == mag1c.NODATA).all()                                      # <=10 px group skipped
    assert np.abs(mf.numpy() - ref).max() <= 2e-3 * np.abs(ref).max()
    assert np.allclose(al.numpy(), g["fbg_albedo"], rtol=1e-5)


def test_features(golden):
    g = golden("features.npz")
    r = features.ratio_2c_match_c_from_sums_outlier(g["bg"].copy(), g["sig"].copy())
    assert np.array_equal(r, g["ratio"])
    assert np.array_equal(features.weight_mag1c(g["mag1c"]), g["weight_mag1c"])
    m = features.ratio_mlr_local(list(g["mlr_bands"]), g["mlr_target"].copy())
    assert np.allclose(m, g["mlr_ratio"], rtol=1e-4, atol=1e-5)


def test_tiling_matches_notebook_counts():
    # (bonus)_training_demo.ipynb:580: 441 chips of 128x128 from 9 tiles
    wins = tiling.create_windows((512, 512), (128, 128), (64, 64))
    assert len(wins) == 49 and 9 * len(wins) == 441
    assert wins[0] == (0, 0, 128, 128) and wins[-1] == (384, 384, 128, 128)
    assert tiling.create_windows((512, 512), (512, 512), (0, 0)) == [(0, 0, 512, 512)]
    assert tiling.tile_id("ang2019", wins[8]) == "ang2019_r64_c64_w128_h128"
    lab = np.zeros((128, 128)); lab[:5, :8] = 1          # 40 px == 10/64^2 * 128^2 -> not >
    assert not tiling.has_plume(lab)
    lab[5, 0] = 1
    assert tiling.has_plume(lab)
    assert tiling.find_padding(1242) == (3, 3) and tiling.find_padding(1280) == (0, 0)


def test_binary_opening_against_scipy_interior():
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(0)
    x = rng.random((2, 1, 40, 40)) > 0.35
    got = morphology.binary_opening(torch.from_numpy(x)).numpy()
    st = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], bool)
    for b in range(2):
        # geodesic border: out-of-image pixels never erode -> scipy erosion with border_value=1
        er = ndi.binary_erosion(x[b, 0], st, border_value=1)
        ref = ndi.binary_dilation(er, st, border_value=0)
        assert np.array_equal(got[b, 0], ref)
