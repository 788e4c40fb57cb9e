"""SURVEY 8(f) rows on the GPU: threshold baselines + validation report (f-4), SRF band aggregation (f-3), training
augmentation + data module + fit loop / checkpoints (f-1), mag1c_emit + EMIT whole-scene / tiled inference (f-2)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from starcop_b200 import augment, baselines, emit, mag1c, srf, synthetic, tiling, trainer, validation  # noqa: E402
from starcop_b200.datamodule import get_dataset  # noqa: E402
from starcop_b200.model_setup import get_model  # noqa: E402
from starcop_b200.settings import default_settings  # noqa: E402

DEV = "cuda"


def to_dev(b):
    return {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in b.items()}


# ---------------------------------------------------------------------------------------------- f-4
WV3 = ["ratio_wv3_B7_B5_varon21_sum_c_out", "ratio_wv3_B8_B8MLR_SanchezGarcia22_sum_c_out", "mag1c"]


def _baseline_batch(seed, B=3, S=96):
    rng = np.random.default_rng(seed)
    b = synthetic.hyperstarcop_batch(B, size=S, seed=seed, channels=3)
    x = b["input"].clone()
    lab = b["output"][:, 0].numpy().astype(bool)
    for c in (0, 1):                                     # ratio bands: small noise, positive inside the plumes
        r = rng.normal(0.0, 0.004, size=lab.shape).astype(np.float32)
        r[lab] += rng.uniform(0.0, 0.02, size=int(lab.sum())).astype(np.float32)
        x[:, c] = torch.from_numpy(r)
    x[:, 2] = b["input"][:, 0]                           # mag1c
    b["input"] = x
    return b


@pytest.mark.parametrize("which", ["mag1c", "sanchez", "varon", "varon_raw_noopen"])
def test_threshold_baselines_match_oracle(which):
    from oracle import baselines as ob
    mk = {"mag1c": (lambda m: m.Mag1cBaseline(WV3)), "sanchez": (lambda m: m.SanchezBaseline(WV3)),
          "varon": (lambda m: m.VaronBaseline(WV3)),
          "varon_raw_noopen": (lambda m: m.VaronBaseline(WV3, 0.004, use_normalisation=False, use_morphological_ops=False))}[which]
    mine, orc = mk(baselines).to(DEV), mk(ob)
    batch = _baseline_batch(3)
    bm, bo = mine.batch_with_preds(to_dev(batch)), orc.batch_with_preds(batch)
    for k in ("pred_binary", "differences", "pred_classification"):
        assert torch.equal(bm[k].cpu(), bo[k]), k                       # masks: bit exact
    assert torch.equal(bm["prediction"].cpu(), bo["prediction"])
    assert torch.equal(bm["input_norm"].cpu(), bo["input_norm"])
    assert bo["pred_binary"].sum() > 0
    for thr in (0.0, 0.01, 500.0):
        assert torch.equal(mine.apply_threshold(bm["prediction"], thr).cpu(), orc.apply_threshold(bo["prediction"], thr))


def test_validation_report_with_baseline_and_model_matches_oracle():
    """run_validation + aggregate (validation.py:80-222) for a threshold baseline (sweep through apply_threshold)
    and for the U-Net module (sweep through the one-pass histogram kernel) against the oracle's pandas aggregation."""
    from oracle import baselines as ob
    from oracle import loss_metrics as olm
    from oracle import validation as ov
    tiles = []
    for i in range(6):
        b = _baseline_batch(20 + i, B=1, S=128)
        if i == 4:
            b["output"].zero_()                                          # a plume-free tile: the (False, "hard") group
        b["has_plume"] = (b["output"].sum((1, 2, 3)) > 0).long()
        tiles.append(b)
    tiles[0]["output"][0, 0, :40, :40] = 1.0                             # > 1000 label pixels: an "easy" tile
    mine, orc = baselines.Mag1cBaseline(WV3).to(DEV), ob.Mag1cBaseline(WV3)
    rows, gcm, sweep = validation.run_validation(mine, tiles)
    # oracle rows
    orows, ocm = [], torch.zeros(2, 2, dtype=torch.long)
    for t in tiles:
        bo = orc.batch_with_preds(t)
        cm = olm.confusion_matrix(bo["pred_binary"], bo["output_norm"].long())
        ocm += cm
        orows.append({"id": t["id"][0], "TP": int(cm[1, 1]), "TN": int(cm[0, 0]), "FP": int(cm[0, 1]), "FN": int(cm[1, 0]),
                      "label_pixels_plume": int(t["output"].sum()), "pred_classification": int(bo["pred_classification"][0, 0])})
    assert torch.equal(gcm, ocm)
    for r, o in zip(rows, orows):
        for k in o:
            assert r[k] == o[k], (k, r[k], o[k])
    # sweep through apply_threshold: exact against the oracle's opening at every threshold
    for thr, cm in sweep:
        ref = torch.zeros(2, 2, dtype=torch.long)
        for t in tiles:
            ref += olm.confusion_matrix(orc.apply_threshold(t["input"][:, 2:3], thr), t["output"].long())
        assert torch.equal(cm, ref), thr
    rows2, rep = validation.aggregate(rows, gcm, sweep)
    _, orep = ov.aggregate(orows, ocm, None)
    for k, v in orep.items():
        if torch.is_tensor(v):
            assert torch.equal(rep[k], v), k
        else:
            assert (np.isnan(v) and np.isnan(rep[k])) or abs(rep[k] - v) <= 1e-12 * max(1.0, abs(v)), (k, rep[k], v)
    assert len(rep["thresholded"]) == 16 and rep["thresholded"][0]["threshold"] > rep["thresholded"][-1]["threshold"]
    # the U-Net module: fused one-pass sweep == per-threshold bincount
    torch.manual_seed(3)
    model = get_model(default_settings(pos_weight=1.0), None).to(DEV).eval()
    mt = [synthetic.hyperstarcop_batch(1, size=64, seed=70 + i) for i in range(3)]
    _, gcm2, sweep2 = validation.run_validation(model, mt)
    for thr, cm in sweep2:
        ref = torch.zeros(2, 2, dtype=torch.long)
        for t in mt:
            with torch.no_grad():
                b = model.batch_with_preds(to_dev(t))
            ref += olm.confusion_matrix((b["prediction"] > thr).long().cpu(), b["output_norm"].long().cpu())
        assert torch.equal(cm, ref), thr
    assert int(gcm2.sum()) == 3 * 64 * 64


# ---------------------------------------------------------------------------------------------- f-3
@pytest.mark.parametrize("H,W,C,K", [(32, 48, 125, 8), (17, 23, 125, 3), (64, 64, 285, 5), (8, 8, 50, 1)])
def test_srf_aggregation_vs_oracle(H, W, C, K):
    import pandas as pd
    from oracle import srf as osrf
    rng = np.random.default_rng(H * 31 + C)
    centers = 380.0 + (2120.0 / C) * np.arange(C)
    wl = np.arange(centers[0] + 1, centers[-1] - 1, 1.0)
    resp = np.stack([np.exp(-0.5 * ((wl - (500 + 1500 * (k + 0.5) / K)) / (12 + 6 * k)) ** 2) for k in range(K)])
    Wt = srf.srf_weight_table(wl, resp, centers)
    cube = rng.uniform(0.5, 9.0, size=(H, W, C)).astype(np.float32)
    cube[H // 2, W // 3, :] = 0.0                        # a nodata pixel: every band equals the fill value
    cube[1, 1, int(np.nonzero(Wt[0])[0][0])] = 0.0       # one band of output 0's window missing
    df = pd.DataFrame(resp.T, index=wl, columns=[f"B{k}" for k in range(K)])
    ref = osrf.transform_to_srf(np.ascontiguousarray(cube.transpose(2, 0, 1)), list(df.columns), df, centers, 0.0)
    out = srf.transform_to_srf(torch.from_numpy(cube).to(DEV), Wt, 0.0).cpu().numpy()
    assert out.shape == (K, H, W)
    assert np.array_equal(out == 0.0, ref == 0.0)                        # missing-value rule: exact
    assert np.allclose(out, ref, rtol=2e-6, atol=1e-6)
    # batched tiles + linearity (size-independent property): srf(a x + b y) = a srf(x) + b srf(y) away from nodata
    x2 = torch.from_numpy(rng.uniform(1, 2, size=(2, H, W, C)).astype(np.float32)).to(DEV)
    a = srf.transform_to_srf(x2, Wt, 0.0)
    b = srf.transform_to_srf(2.0 * x2, Wt, 0.0)
    assert a.shape == (2, K, H, W) and torch.allclose(b, 2.0 * a, rtol=1e-6)


def test_srf_full_size_rows_sum_to_one():
    """512 x 512 x 125 (BASELINE.json configs[2] cube shape): a constant cube maps to the same constant in every
    output band (weights sum to one) -- a size-independent check at the bench size."""
    C, K = 125, 8
    centers = 380.0 + 17.0 * np.arange(C)
    wl = np.arange(400.0, 2400.0, 2.0)
    resp = np.stack([np.exp(-0.5 * ((wl - (450 + 230 * k)) / 40.0) ** 2) for k in range(K)])
    Wt = srf.srf_weight_table(wl, resp, centers)
    cube = torch.full((2, 512, 512, C), 3.25, device=DEV)
    out = srf.transform_to_srf(cube, Wt, 0.0)
    assert torch.allclose(out, torch.full_like(out, 3.25), rtol=1e-6)


# ---------------------------------------------------------------------------------------------- f-1
def test_affine_warp_matches_oracle():
    from oracle import augment as oa
    g = torch.Generator().manual_seed(5)
    B, C, H, W = 6, 3, 40, 56
    x = torch.rand(B, C, H, W, generator=g)
    params = augment.draw_params(B, torch.Generator().manual_seed(9))
    params["angle"][0] = 0.0; params["hflip"][0] = True; params["vflip"][0] = False      # a pure flip: exact permutation
    mats = augment.dst_to_src_matrices(params, H, W)
    for nearest in (False, True):
        out = augment.affine_warp(x.to(DEV), mats, nearest=nearest).cpu()
        for i in range(B):
            ref = oa.augment_sample(x[i], float(params["angle"][i]), bool(params["hflip"][i]), bool(params["vflip"][i]),
                                    mode="nearest" if nearest else "bilinear")
            if nearest:       # rounding ties at exactly .5 may pick the other neighbour: compare where both agree on geometry
                assert (out[i] != ref).float().mean().item() < 0.01
            else:
                assert (out[i] - ref).abs().max().item() <= 2e-5, i
    assert torch.equal(augment.affine_warp(x.to(DEV), mats)[0].cpu(), torch.flip(x[0], dims=(-1,)))
    # a reproducible stream: the same seed gives the same augmented batch
    b = to_dev(synthetic.hyperstarcop_batch(4, size=64, seed=1))
    a1, a2 = augment.TrainAugmentation(seed=3)(b), augment.TrainAugmentation(seed=3)(b)
    assert torch.equal(a1["input"], a2["input"]) and torch.equal(a1["output"], a2["output"])
    lab = augment.TrainAugmentation(seed=3, mask_mode="nearest")(b)["output"]
    assert set(torch.unique(lab).tolist()) <= {0.0, 1.0}


def test_fit_loop_checkpoint_resume_and_scheduler(tmp_path):
    """trainer.Trainer = the reference's Trainer.fit order: augmented chips from the weighted sampler, validation at
    val_check_interval on full scenes, ReduceLROnPlateau -> device lr, top-1 checkpoint by val_loss, final
    checkpoint; a checkpoint restores weights + Adam state bit-exactly and loads into a fresh ModelModule."""
    st = default_settings(pos_weight=1.0, compute_dtype="bf16")
    st.dataloader.batch_size, st.dataloader.num_workers = 16, 0
    st.dataset.training_size, st.dataset.training_size_overlap = [64, 64], [0, 0]
    st.training.max_epochs, st.training.val_check_interval = 2, 0.5
    st.seed = 7
    dm = get_dataset(st, n_train_scenes=2, n_test_scenes=2, scene_size=256)
    logs = []
    model, tr, report = trainer.train(st, data_module=dm, experiment_path=str(tmp_path), log=logs.append)
    assert len(dm.train_dataset) == 2 * 16                                # 256 / 64 = 4 x 4 windows per scene
    assert len(tr.history) == 4 and tr.global_step == 2 * 2               # 2 epochs x 2 batches, validation every batch
    assert all(np.isfinite(h["val_loss"]) for h in tr.history) and "val_iou" in tr.history[0]
    assert tr.history[-1]["val_loss"] < tr.history[0]["val_loss"] * 1.5
    assert os.path.exists(tr.best_path) and len(os.listdir(os.path.join(tmp_path, "checkpoint"))) == 1     # save_top_k = 1
    assert tr.best_score == min(h["val_loss"] for h in tr.history)
    final = os.path.join(tmp_path, "final_checkpoint_model.ckpt")
    ck = torch.load(final, map_location="cpu", weights_only=False)
    assert ck["global_step"] == 4 and "network.encoder.features.0.0.weight" in ck["state_dict"]
    assert "f1score" in report and os.path.exists(os.path.join(tmp_path, "results_agg.json"))
    # load_from_checkpoint surface + bit-exact weights
    from starcop_b200.model_module import ModelModule
    m2 = ModelModule.load_from_checkpoint(final, settings=st).to(DEV)
    for (n, a), (_, b) in zip(model.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a.cpu(), b.cpu()), n
    # resume: optimiser moments, step counter and lr come back
    tr2 = trainer.Trainer(max_epochs=2, checkpoint_dir=None, use_cuda_graph=False, log=lambda *_: None)
    cfg = m2.configure_optimizers()
    tr2._resume(m2, cfg["lr_scheduler"], final)
    s1, s2 = model.network._adam_state, m2.network._adam_state
    assert torch.equal(s1["m"], s2["m"]) and torch.equal(s1["v"], s2["v"]) and int(s1["step"]) == int(s2["step"])
    b = to_dev(synthetic.hyperstarcop_batch(4, size=64, seed=2))
    l1, l2 = model.train_step_fused(b), m2.train_step_fused(b)
    assert l1.item() == l2.item() and torch.equal(model.network.flat_params, m2.network.flat_params)


# ---------------------------------------------------------------------------------------------- f-2
def test_mag1c_emit_column_groups_vs_oracle():
    from oracle import mag1c as om
    rng = np.random.default_rng(4)
    rows, cols, bands = 96, 12, 64
    wl = np.linspace(1900.0, 2500.0, bands)
    sel = (wl >= 2122) & (wl <= 2488)
    S = int(sel.sum())
    t = synthetic.synthetic_template(S)
    mu = 6.0 * np.exp(-np.arange(bands) / 50.0) + 0.5
    albedo = rng.uniform(0.5, 1.5, size=(rows, cols, 1))
    alpha = np.zeros((rows, cols, 1)); alpha[30:50, 4:8] = 0.02
    tfull = np.zeros(bands); tfull[sel] = t
    raw = (albedo * mu * (1 + alpha * tfull) + rng.normal(0, 0.01, size=(rows, cols, bands)) * mu).astype(np.float32)
    raw[5, 3, :] = -9999.0                               # an invalid pixel
    raw[:, 10:12, :] = -9999.0                           # a fully invalid column group (skipped: stays fill)
    for step in (2, 5):
        mf_o, al_o = om.mag1c_emit(raw, wl, t, column_step=step, num_iter=30)
        mf, al = mag1c.mag1c_emit(torch.from_numpy(raw).to(DEV), wl, template=t, column_step=step, num_iter=30)
        mf, al = mf.cpu().numpy(), al.cpu().numpy()
        assert np.array_equal(mf == -9999.0, mf_o == -9999.0)            # fill / invalid bookkeeping: exact
        ok = mf_o != -9999.0
        scale = np.abs(mf_o[ok]).max()
        assert np.abs(mf[ok] - mf_o[ok]).max() <= 1e-5 * scale, (step, np.abs(mf[ok] - mf_o[ok]).max(), scale)
        assert np.allclose(al[ok], al_o[ok], rtol=1e-6)
    assert mf[30:50, 4:8].mean() > 5 * np.abs(mf[60:, :8]).mean()        # the injected plume is recovered


def test_emit_scene_prediction_tiled_equals_whole_scene_semantics():
    """predict_scene: the un-tiled pass IS padded_predict (oracle-checked in test_gpu_unet); tiles of size T give
    the same probabilities as running the model on each reflect-padded tile separately (stitching exact)."""
    torch.manual_seed(2)
    model = get_model(default_settings(pos_weight=1.0), None).to(DEV).eval()
    scene = torch.from_numpy(synthetic.hyperstarcop_batch(1, size=160, seed=9)["input"][0][:, :150, :139].numpy()).to(DEV)
    whole = emit.predict_scene(scene, model)
    ref = torch.sigmoid(torch.from_numpy(tiling.padded_predict(scene, model))).to(DEV)
    assert whole.shape == (1, 150, 139) and torch.allclose(whole, ref, atol=1e-6)
    T = 64
    tiled = emit.predict_scene(scene, model, tile=T, batch=4)
    pr, pc = tiling.find_padding(150, T), tiling.find_padding(139, T)
    padded = torch.nn.functional.pad(scene[None], (pc[0], pc[1], pr[0], pr[1]), mode="reflect")
    man = torch.zeros(1, padded.shape[-2], padded.shape[-1], device=DEV)
    with torch.no_grad():
        for i in range(0, padded.shape[-2], T):
            for j in range(0, padded.shape[-1], T):
                man[:, i:i + T, j:j + T] = torch.sigmoid(model(padded[:, :, i:i + T, j:j + T]))[0]
    assert torch.allclose(tiled, man[:, pr[0]:pr[0] + 150, pc[0]:pc[0] + 139], atol=2e-6)
    # the CUDA-graph replay (default) and the launch-by-launch path run the same kernels on the same addresses
    assert torch.equal(tiled, emit.predict_scene(scene, model, tile=T, batch=4, graphed=False))
    assert torch.equal(whole, emit.predict_scene(scene, model, graphed=False))
    assert torch.equal(tiled, emit.predict_scene(scene, model, tile=T, batch=4))          # second replay of the cached graph
    # full EMIT front end on a small cube: mag1c_emit -> RGB pick -> rescale -> (4, H32, W32)
    rng = np.random.default_rng(1)
    wl = np.linspace(400.0, 2500.0, 120)
    raw = torch.from_numpy((rng.uniform(0.5, 1.5, size=(70, 40, 1)) * (4.0 * np.exp(-np.arange(120) / 80.0) + 0.5)
                            * (1 + rng.normal(0, 0.01, size=(70, 40, 120)))).astype(np.float32)).to(DEV)
    S = int(((wl >= 2122) & (wl <= 2488)).sum())
    x, mf, al = emit.emit_model_input(raw, wl, template=synthetic.synthetic_template(S), column_step=4)
    assert x.shape == (4, 64, 32) and torch.isfinite(x).all() and float(x[0].max()) <= 3500.0 and float(x[1:].max()) <= 120.0
    assert emit.rgb_band_indices(wl) == [int(np.argmin(np.abs(wl - w))) for w in (640, 550, 460)]


def test_cube_to_unet_chain_fused_pack_equals_the_dataloader_contract():
    """configs[2]: the fused pack (producer writing the engine's NHWC input) gives the SAME logits / loss as feeding
    the materialised (B,4,H,W) batch [clip(mf,0,1e4), R, G, B] through normalize_x, and weight_loss = weight_mag1c."""
    from starcop_b200 import chain, features
    cube_np, t73, _ = synthetic.aviris_cube(2, size=64, bands=125, seed=3)
    cube = torch.from_numpy(cube_np).to(DEV)
    sl = slice(52, 125)
    rgb = chain.rgb_bands(380.0 + 5.0 * np.arange(125))
    y = (torch.rand(2, 1, 64, 64, device=DEV) > 0.9).float()
    for mode in ("f32", "bf16"):
        torch.manual_seed(4)
        m = get_model(default_settings(pos_weight=1.0, compute_dtype=mode), None).to(DEV).eval()
        a = chain.cube_batch(m, cube, t73, sl, rgb, output=y, materialize_input=True)
        b = chain.cube_batch(m, cube, t73, sl, rgb, output=y)
        mf, _ = mag1c.mag1c_tiles(cube, t73, sl, num_iter=30)
        assert torch.equal(a["input"][:, 0], mf.clamp(0, 10000))
        for k, bi in enumerate(rgb):
            assert torch.equal(a["input"][:, 1 + k], cube[..., bi])
        assert torch.equal(a["weight_loss"], features.weight_mag1c(a["input"][:, 0:1]))
        with torch.no_grad():
            la = m(a["input"])
            lb = m.network._forward_impl(b["input"], None, False, record=False)
        assert torch.equal(la, lb)
        m.train()
        torch.manual_seed(4)
        m2 = get_model(default_settings(pos_weight=1.0, compute_dtype=mode), None).to(DEV).train()
        l1, l2 = m.train_step_fused(a), m2.train_step_fused(b)
        assert l1.item() == l2.item() and torch.equal(m.network.flat_params, m2.network.flat_params)
