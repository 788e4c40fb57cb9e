"""CPU-only checks: the C-ABI library loads and exports every declared symbol, host logic."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from starcop_b200 import _lib, build
    build.build()
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "starcop_b200.h")).read()
    declared = set(re.findall(r"\b(sc_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/starcop_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.sc_abi_version() == 1


def test_fastdiv_constants_are_exact():
    """the multiply-high division the persistent tile loops use to decode (image, tile row, tile column): exact for
    every divisor / dividend class the kernels can see (0 <= x < 2^31), evaluated by the library's own host code"""
    import random
    from starcop_b200 import _lib, build
    build.build()
    lib = _lib.load()
    rng = random.Random(0)
    divisors = list(range(1, 70)) + [127, 128, 129, 255, 256, 257, 1000, 4095, 4096, 4097, 65535, 65536, 1 << 20, (1 << 24) + 3]
    for d in divisors:
        xs = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, (1 << 31) - 1, (1 << 31) - 2] + [rng.randrange(1 << 31) for _ in range(200)]
        for x in xs:
            assert lib.sc_debug_fastdiv(x, d) == x // d, (x, d)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "starcop_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_no_cpu_path():
    from starcop_b200 import _lib
    from starcop_b200.model_setup import get_model
    from starcop_b200.settings import default_settings
    m = get_model(default_settings(), None)
    with pytest.raises(_lib.StarcopB200Error):
        m(torch.zeros(1, 4, 32, 32))


def test_state_dict_keys_and_init_match_oracle():
    from oracle.module import get_model as og
    from starcop_b200.model_setup import get_model
    from starcop_b200.settings import default_settings
    torch.manual_seed(7); m = get_model(default_settings(), None)
    torch.manual_seed(7); o = og(default_settings())
    sm, so = m.network.state_dict(), o.network.state_dict()
    assert list(sm) == list(so)
    assert all(torch.equal(sm[k], so[k]) for k in sm)
    extra = [k for k in m.state_dict() if not k.startswith("network.")]
    assert set(extra) == {"pos_weight", "loss_function.pos_weight", "normalizer.offsets_input",
                          "normalizer.factors_input", "normalizer.clip_min_input", "normalizer.clip_max_input"}
    assert m.normalizer.factors_input.dtype == torch.int64        # python ints -> int64


def test_metrics_match_reference_golden(golden):
    from starcop_b200 import metrics
    g = golden("metrics.npz")
    for cm, vals in zip(g["cms"], g["values"]):
        for n, v in zip(g["names"], vals):
            got = float(getattr(metrics, str(n))(torch.from_numpy(cm)))
            assert (np.isnan(got) and np.isnan(v)) or got == v, (cm, n, got, v)


def test_synthetic_batch_contract():
    from starcop_b200 import synthetic
    b = synthetic.hyperstarcop_batch(4, size=64, seed=0)
    assert b["input"].shape == (4, 4, 64, 64) and b["input"].dtype == torch.float32
    assert b["output"].shape == (4, 1, 64, 64) and set(b["output"].unique().tolist()) <= {0.0, 1.0}
    assert b["weight_loss"].min() >= 0.1 and b["weight_loss"].max() <= 1.0
    assert b["has_plume"].dtype == torch.int64 and len(b["id"]) == 4


def test_tiling_matches_oracle_and_notebook_counts():
    from oracle import tiling as ot
    from starcop_b200 import tiling
    for args in (((512, 512), (128, 128), (64, 64)), ((512, 512), (512, 512), (0, 0)), ((300, 500), (128, 128), (32, 64))):
        assert tiling.create_windows(*args) == ot.create_windows(*args)
    wins = tiling.create_windows((512, 512), (128, 128), (64, 64))
    assert len(wins) == 49                                     # 441 chips / 9 tiles in the reference notebook
    assert tiling.tile_id("x", wins[1]) == ot.tile_id("x", wins[1]) == "x_r0_c64_w128_h128"
    lab = np.zeros((128, 128), np.float32); lab[:5, :8] = 1
    assert tiling.has_plume(torch.from_numpy(lab)) == ot.has_plume(lab) == False
    for v in (1242, 1280, 70, 8):
        assert tiling.find_padding(v, 32) == ot.find_padding(v, 32)


def test_generate_template_known_answer_on_a_synthetic_lut():
    """A9 (mag1c.py:60-95) on an analytic look-up table: radiance = exp(-k(lambda) * c) has log-slope -k, so the
    template of a narrow band is -k(center) * 1e5; also the reference's argument checks."""
    import numpy as np
    import pytest
    from starcop_b200 import mag1c
    wave = np.linspace(2000.0, 2500.0, 5001)
    k = 1e-5 * (1.0 + 0.5 * np.sin(wave / 40.0))
    rads = np.exp(-k[None, :] * np.asarray(mag1c.CH4_CONCENTRATIONS, dtype=np.float64)[:, None])
    centers = np.array([2100.0, 2250.0, 2400.0])
    out = mag1c.generate_template_from_bands(centers, np.full(3, 0.5), lut=(rads, wave))
    assert out.shape == (3, 2) and np.array_equal(out[:, 0], centers)
    assert np.allclose(out[:, 1], -np.interp(centers, wave, k) * 1e5, rtol=2e-4)
    # a wide band averages the RADIANCE (not k): compare with the definition evaluated directly
    wide = mag1c.generate_template_from_bands([2250.0], [30.0], lut=(rads, wave))[0, 1]
    g = np.exp(-(wave - 2250.0) ** 2 / (2 * (30.0 / 2.3548200450309493) ** 2)); g /= g.sum()
    slope = np.polyfit(np.asarray(mag1c.CH4_CONCENTRATIONS, float), np.log(rads @ g), 1)[0]
    assert abs(wide - slope * 1e5) <= 1e-9 * abs(wide)
    with pytest.raises(RuntimeError):
        mag1c.generate_template_from_bands([2100.0, np.nan], [5.0, 5.0], lut=(rads, wave))
    with pytest.raises(RuntimeError):
        mag1c.generate_template_from_bands([2100.0, 2200.0], [5.0], lut=(rads, wave))


def test_generate_template_matches_reference_golden_when_the_lut_is_present(golden):
    """with the reference's ch4.hdr / ch4.lut at hand (build container) the 73-band AVIRIS template equals the one
    the reference's own function produced (tests/golden/ch4_template_aviris.npz)"""
    import os
    import numpy as np
    import pytest
    from starcop_b200 import mag1c
    lut_dir = os.environ.get("STARCOP_CH4_LUT_DIR", "/root/reference/starcop/models")
    if not os.path.exists(os.path.join(lut_dir, "ch4.lut")):
        pytest.skip("ch4.lut not available on this machine")
    g = golden("ch4_template_aviris.npz")
    rads, wave = mag1c.read_ch4_lut(lut_dir)
    assert rads.shape == (7, 31800) and wave.shape == (31800,)
    out = mag1c.generate_template_from_bands(g["centers"], g["fwhm"], lut_dir=lut_dir)
    assert np.allclose(out, g["template"], rtol=1e-10, atol=0)
    sl = mag1c.band_keep_aviris(g["wavelengths"])
    assert (sl.start, sl.stop - 1) == (int(g["band_first"]), int(g["band_last"]))


def test_tiled_records_and_sample_weights():
    """A15: datamodule.py:17-64 (tiling of the 512 x 512 scenes, has_plume, id format) and :309-315 (sampler weights)"""
    import numpy as np
    from starcop_b200 import tiling
    lab = np.zeros((2, 512, 512), np.float32)
    lab[0, 64:128, 64:128] = 1                     # 4096 positives, inside tiles r0/r64 x c0/c64
    lab[1, 0:3, 0:3] = 1                           # 9 px: 9/128^2 < 10/64^2 -> no plume anywhere
    recs = [{"id": "ang1", "window_row_off": 0, "window_col_off": 0, "window_width": 512, "window_height": 512, "qplume": 1.0},
            {"id": "ang2", "window_row_off": 0, "window_col_off": 0, "window_width": 512, "window_height": 512, "qplume": 0.0}]
    tiles = tiling.tiled_records(recs, lab, (128, 128), (64, 64))
    assert len(tiles) == 2 * 49                    # 7 x 7 windows per scene (the notebook's 441 = 9 x 49)
    t = {x["id"]: x for x in tiles}
    assert "ang1_r0_c0_w128_h128" in t and t["ang1_r64_c64_w128_h128"]["id_original"] == "ang1"
    a = t["ang1_r64_c64_w128_h128"]
    assert a["frac_positives"] == 4096 / 128 ** 2 and a["has_plume"] is True and a["qplume"] == 1.0
    assert (a["window_row_off"], a["window_col_off"], a["window_width"], a["window_height"]) == (64, 64, 128, 128)
    assert t["ang1_r0_c0_w128_h128"]["frac_positives"] == 4096 / 128 ** 2       # the blob sits in all four overlapping tiles
    assert t["ang1_r128_c128_w128_h128"]["has_plume"] is False
    assert not any(x["has_plume"] for x in tiles if x["id_original"] == "ang2")
    flags = np.array([x["has_plume"] for x in tiles])
    n_plume = int(flags.sum())
    assert n_plume == 4
    w = tiling.add_sample_weight(flags)
    assert np.allclose(w[flags], len(tiles) / n_plume) and np.allclose(w[~flags], len(tiles) / (len(tiles) - n_plume))
    assert abs(w[flags].sum() - w[~flags].sum()) < 1e-9      # both classes get the same total sampling mass
    # whole-tile mode (tile = scene): one window, has_plume of the full label
    whole = tiling.tiled_records(recs, lab, (512, 512), (0, 0))
    assert [x["id"] for x in whole] == ["ang1_r0_c0_w512_h512", "ang2_r0_c0_w512_h512"]
    assert whole[0]["has_plume"] is True and whole[1]["has_plume"] is False


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the reference's CPU path, oracle port) runs without a GPU and prints exactly one
    JSON line on stdout carrying the contract's keys"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--size", "64"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tiles/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k


def test_load_settings_yaml_and_hydra_style_overrides(tmp_path):
    """scripts/train.py's settings without hydra: the reference's own config.yaml (when at hand) or a copy of its
    structure, plus key=value overrides as in bash/bash_train_example.sh"""
    import os
    import pytest
    from starcop_b200 import settings as st
    ref = "/root/reference/scripts/configs/config.yaml"
    if os.path.exists(ref):
        s = st.load_settings(ref, ["model.pos_weight=1", "experiment_name=HyperSTARCOP_magic_rgb",
                                   "dataset.input_products=[mag1c,TOA_AVIRIS_640nm,TOA_AVIRIS_550nm,TOA_AVIRIS_460nm]"])
        assert s.model.pos_weight == 1 and s.model.lr == 0.0001 and s.model.semseg_backbone == "mobilenet_v2"
        assert s.dataset.input_products == st.HYPERSTARCOP_PRODUCTS and s.dataset.use_weight_loss is True
        assert "use_weight_loss" in s.dataset and "hydra" not in s and s.dataloader.batch_size == 32
        assert s.dataset.training_size == [128, 128] and s.experiment_name == "HyperSTARCOP_magic_rgb"
    y = tmp_path / "c.yaml"
    y.write_text("model:\n  lr: 0.001\n  pos_weight: 15\ndataset:\n  input_products: [a, b]\nhydra:\n  job: {chdir: true}\n")
    s = st.load_settings(str(y), ["model.lr=1e-4", "+model.compute_dtype=bf16", "dataset.input_products=[mag1c]"])
    assert s.model.lr == 1e-4 and s.model.compute_dtype == "bf16" and s.dataset.input_products == ["mag1c"]
    assert "hydra" not in s
    with pytest.raises(KeyError):
        st.load_settings(str(y), ["model.nope=1"])
    with pytest.raises(ValueError):
        st.load_settings(str(y), ["model.lr"])
    d = st.load_settings(None, ["model.pos_weight=15"])
    assert d.model.pos_weight == 15 and d.dataset.input_products == st.HYPERSTARCOP_PRODUCTS


# ---- SURVEY 8(f) host logic (no GPU) -------------------------------------------------------------------------------
def test_synthetic_datamodule_contract_and_chip_count():
    """Permian2019DataModule surface over synthetic scenes: 9 scenes x 49 windows = 441 chips (the reference
    notebook's printout), the STARCOPDataset item dict, weighted sampling over add_sample_weight."""
    import torch
    from starcop_b200 import tiling
    from starcop_b200.datamodule import get_dataset
    from starcop_b200.settings import default_settings
    st = default_settings()
    st.dataloader.num_workers = 0
    dm = get_dataset(st, n_train_scenes=9, n_test_scenes=2)
    dm.prepare_data()
    assert len(dm.train_dataset) == 441 and len(dm.val_dataset) == 2 and len(dm.train_dataset_non_tiled) == 9
    it = dm.train_dataset[0]
    assert set(it) == {"input", "output", "weight_loss", "id", "has_plume"}
    assert it["input"].shape == (4, 128, 128) and it["output"].shape == (1, 128, 128) and it["input"].dtype == torch.float32
    assert it["id"].endswith("_r0_c0_w128_h128") and isinstance(it["has_plume"], int)
    rec = dm.train_dataset.records[10]
    r, c, h, w = rec["window"]
    assert (r, c, h, w) == (64, 192, 128, 128) and rec["id"].endswith(f"_r{r}_c{c}_w{w}_h{h}")
    scene = dm.train_dataset.scenes[rec["scene"]]
    assert rec["has_plume"] == (float(scene["output"][:, r:r + h, c:c + w].sum()) / (h * w) > 10 / 64 ** 2)
    g = torch.Generator().manual_seed(0)
    b = next(iter(dm.train_dataloader(batch_size=8, generator=g)))
    assert b["input"].shape == (8, 4, 128, 128) and b["has_plume"].shape == (8,) and len(b["id"]) == 8
    # the sampler favours the rarer class like the reference's 1/fraction weights
    flags = [r["has_plume"] for r in dm.train_dataset.records]
    w = tiling.add_sample_weight(flags)
    frac = sum(flags) / len(flags)
    assert abs(w[flags.index(True)] - 1 / frac) < 1e-12 and abs(w[flags.index(False)] - 1 / (1 - frac)) < 1e-12
    v = next(iter(dm.val_dataloader(batch_size=1)))
    assert v["input"].shape == (1, 4, 512, 512)


def test_validation_aggregate_matches_reference_pandas_path():
    import torch
    from oracle import validation as ov
    from starcop_b200 import validation
    rng = np.random.default_rng(0)
    rows = []
    for i in range(12):
        lab = int(rng.choice([0, 0, 300, 5000]))
        tp = int(rng.integers(0, lab + 1)) if lab else 0
        fp = int(rng.integers(0, 400))
        rows.append({"id": f"t{i}", "TP": tp, "FN": lab - tp, "FP": fp, "TN": 128 * 128 - lab - fp, "label_pixels_plume": lab,
                     "pred_classification": int(tp + fp > 40), "has_plume": int(lab > 0)})
    rows[0].update(label_pixels_plume=0, TP=0, FN=0, TN=128 * 128 - rows[0]["FP"])      # the groups the reference indexes exist
    rows[1].update(label_pixels_plume=300, TP=100, FN=200, TN=128 * 128 - 300 - rows[1]["FP"])
    rows[2].update(label_pixels_plume=5000, TP=4000, FN=1000, TN=128 * 128 - 5000 - rows[2]["FP"])
    gcm = torch.tensor([[sum(r["TN"] for r in rows), sum(r["FP"] for r in rows)], [sum(r["FN"] for r in rows), sum(r["TP"] for r in rows)]])
    sweep = [(0.9, gcm.clone()), (0.5, gcm.clone())]
    out_rows, rep = validation.aggregate(rows, gcm, sweep)
    _, ref = ov.aggregate(rows, gcm, None)
    for k, v in ref.items():
        if torch.is_tensor(v):
            assert torch.equal(rep[k], v), k
        else:
            assert (np.isnan(v) and np.isnan(rep[k])) or rep[k] == pytest.approx(v, rel=1e-12, abs=0), k
    assert [r["difficulty"] for r in out_rows[:3]] == ["hard", "hard", "easy"]
    assert [d["threshold"] for d in rep["thresholded"]] == [0.9, 0.5] and "FPR" in rep["thresholded"][0]


def test_srf_weight_table_matches_reference_recipe():
    import pandas as pd
    from oracle import srf as osrf
    from starcop_b200 import srf
    rng = np.random.default_rng(1)
    centers = 380.0 + 5.01 * np.arange(125)
    wl = np.arange(400.0, 990.0, 0.5)
    resp = np.stack([np.exp(-0.5 * ((wl - c) / s) ** 2) for c, s in ((450, 15), (560, 20), (665, 12), (842, 45))])
    W = srf.srf_weight_table(wl, resp, centers)
    assert np.allclose(W.sum(1), 1.0, atol=1e-12)
    cube = rng.uniform(1, 5, size=(125, 5, 6)).astype(np.float32)
    df = pd.DataFrame(resp.T, index=wl, columns=list("abcd"))
    ref = osrf.transform_to_srf(cube, list("abcd"), df, centers)
    assert np.allclose(np.einsum("kc,chw->khw", W, cube.astype(np.float64)), ref, rtol=1e-6)
    with pytest.raises(ValueError):
        srf.srf_weight_table(np.array([100.0, 500.0]), np.ones((1, 2)), centers)          # outside the band range: interp1d raises


def test_augmentation_matrices_compose_like_the_reference_operators():
    import torch
    from oracle import augment as oa
    from starcop_b200 import augment
    H, W = 12, 16
    x = torch.rand(2, H, W, generator=torch.Generator().manual_seed(0))
    for ang, hf, vf in [(0, 1, 0), (0, 0, 1), (37.0, 0, 0), (-63.0, 1, 1), (90.0, 1, 0)]:
        p = {"angle": torch.tensor([ang]), "hflip": torch.tensor([bool(hf)]), "vflip": torch.tensor([bool(vf)])}
        m = augment.dst_to_src_matrices(p, H, W)[0].double()
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
        sx, sy = m[0] * xs + m[1] * ys + m[2], m[3] * xs + m[4] * ys + m[5]
        grid = torch.stack([2 * sx / (W - 1) - 1, 2 * sy / (H - 1) - 1], -1)[None].float()
        mine = torch.nn.functional.grid_sample(x[None], grid, mode="bilinear", padding_mode="zeros", align_corners=True)[0]
        assert (mine - oa.augment_sample(x, ang, hf, vf)).abs().max().item() < 5e-6
    d = augment.draw_params(1000, torch.Generator().manual_seed(1))
    assert 0.4 < (d["angle"] != 0).float().mean() < 0.6 and d["angle"].abs().max() <= 90 and 0.4 < d["hflip"].float().mean() < 0.6


def test_train_entry_point_parses_reference_style_overrides():
    from starcop_b200 import settings as S
    ref_cfg = "/root/reference/scripts/configs/config.yaml"
    st = S.load_settings(ref_cfg if os.path.exists(ref_cfg) else None,
                         ["model.pos_weight=1", "training.max_epochs=2", "dataloader.batch_size=16", "+model.compute_dtype=bf16",
                          "dataset.input_products=[mag1c,TOA_AVIRIS_640nm,TOA_AVIRIS_550nm,TOA_AVIRIS_460nm]"])
    assert st.model.pos_weight == 1 and st.training.max_epochs == 2 and st.model.compute_dtype == "bf16"
    assert st.dataset.input_products[0] == "mag1c" and st.training.val_check_interval == 0.5
    import scripts.train as entry
    assert callable(entry.main)
