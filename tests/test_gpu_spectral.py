"""Spectral products on the GPU (mag1c matched filter, band ratio) against the vectors the
reference's own code produced (tests/golden/mag1c.npz, features.npz) and against the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import features as ofeat, mag1c as omag  # noqa: E402
from starcop_b200 import features, mag1c, synthetic  # noqa: E402

DEV = "cuda"


@pytest.mark.parametrize("tag,dt,alpha,tol", [("f32", torch.float32, 0.0, 3e-3), ("f64", torch.float64, 1e-4, 1e-7)])
def test_matched_filter_vs_reference(golden, tag, dt, alpha, tol):
    g = golden("mag1c.npz")
    x = torch.from_numpy(g["x_f32"]).to(dt).to(DEV)
    t = g["template"]
    mf, R = mag1c.rmf(x, t, alpha=alpha)
    ref = g[f"rmf_{tag}_mf"]
    assert np.allclose(R.cpu().numpy(), g[f"rmf_{tag}_R"], rtol=2e-6 if tag == "f32" else 1e-12)
    assert np.abs(mf.cpu().numpy() - ref).max() <= tol * np.abs(ref).max()
    for it in (1, 30):
        mf, R = mag1c.acrwl1mf(x, t, num_iter=it, alpha=alpha)
        ref = g[f"acrwl1mf_{tag}_it{it}_mf"]
        err = np.abs(mf.cpu().numpy() - ref).max()
        assert err <= tol * np.abs(ref).max(), (it, err, np.abs(ref).max())
        assert np.allclose(R.cpu().numpy(), g[f"acrwl1mf_{tag}_it{it}_R"], rtol=2e-6 if tag == "f32" else 1e-12)
        # zeros of the non-negativity constraint agree wherever the reference is clearly positive / zero
        refz = ref == 0
        got = mf.cpu().numpy()
        assert (got[refz] <= tol * np.abs(ref).max()).all()


def test_func_by_groups_matches_reference(golden):
    g = golden("mag1c.npz")
    cube = torch.from_numpy(g["fbg_cube"]).to(DEV)
    mf, al = mag1c.func_by_groups(cube, g["fbg_groups"], g["template"], mask=g["fbg_mask"], num_iter=30)
    ref = g["fbg_mf"]
    got = mf.cpu().numpy()
    assert np.array_equal(got == mag1c.NODATA, ref == mag1c.NODATA)      # skipped groups / masked pixels: bit exact
    assert (ref[:, 10:] == mag1c.NODATA).all()
    assert np.abs(got - ref).max() <= 3e-3 * np.abs(ref).max()
    assert np.allclose(al.cpu().numpy(), g["fbg_albedo"], rtol=1e-5)


def test_tile_columns_straight_from_bip_cube(golden):
    t73 = golden("ch4_template_aviris.npz")["template"][:, 1]
    cube, _, alpha = synthetic.aviris_cube(2, size=128, bands=125, seed=9, template=t73)
    sl = slice(52, 125)
    mf, al = mag1c.mag1c_tiles(torch.from_numpy(cube).to(DEV), t73, sl, num_iter=30)
    for n in range(2):
        # 128-pixel groups x 73 bands: the covariance is barely over-determined, so the reference's
        # fp32 arithmetic is itself ~10 % noisy here.  Yardstick = the same algorithm in float64;
        # the GPU path (fp64 statistics) must be close to it and no worse than the fp32 reference.
        m64, a64 = omag.mag1c_tile_columns(cube[n].astype(np.float64), t73, sl, num_iter=30)
        m32, a32 = omag.mag1c_tile_columns(cube[n], t73, sl, num_iter=30)
        scale = m64.abs().max().item()
        e_gpu = (mf[n].cpu().double() - m64).abs().max().item()
        e_ref = (m32.double() - m64).abs().max().item()
        assert e_gpu <= max(2e-3 * scale, e_ref), (e_gpu, e_ref, scale)
        assert torch.allclose(al[n].cpu().double(), a64, rtol=1e-5)
    # the injected plumes are recovered: the filter output inside them is far above the background
    plume = torch.from_numpy(alpha[0] > 0)
    m = mf[0].cpu()
    assert m[plume].mean() > 5 * m[~plume].mean()


def test_ratio_product_vs_reference(golden):
    g = golden("features.npz")
    bg, sig = torch.from_numpy(g["bg"]).to(DEV), torch.from_numpy(g["sig"]).to(DEV)
    r = features.ratio_2c_match_c_from_sums_outlier(bg, sig).cpu().numpy()
    ref = g["ratio"]
    assert np.array_equal(r == np.float32(-0.6), ref == np.float32(-0.6))     # nodata mask: bit exact
    assert np.allclose(r, ref, rtol=2e-6, atol=2e-6)


def test_ratio_product_batched_tiles_vs_oracle():
    tiles = [synthetic.ratio_bands(96, seed=s) for s in range(3)]
    bg = torch.from_numpy(np.stack([t[0] for t in tiles])).to(DEV)
    sig = torch.from_numpy(np.stack([t[1] for t in tiles])).to(DEV)
    r = features.ratio_2c_match_c_from_sums_outlier(bg[:, None], sig[:, None], p=5).cpu().numpy()[:, 0]
    for i, (b, s) in enumerate(tiles):
        ref = ofeat.ratio_2c_match_c_from_sums_outlier(b.copy(), s.copy())
        assert np.allclose(r[i], ref, rtol=2e-6, atol=2e-6), i
    # exact percentiles: compare the inlier gain against numpy's
    b, s = tiles[0]
    c = ofeat.no_outliers(b.flatten()).sum(dtype=np.float64) / ofeat.no_outliers(s.flatten()).sum(dtype=np.float64)
    ok = (b > 1e-3)
    c_gpu = np.median(((r[0] * (b + 1e-6) + b) / s)[ok])
    assert abs(c_gpu - c) <= 1e-5 * abs(c)


@pytest.mark.parametrize("h,w", [(512, 512), (99, 77), (64, 50), (3, 5)])
def test_ratio_cluster_kernel_matches_single_cta_select_and_numpy(h, w, monkeypatch):
    """the cluster-resident kernels (16 / 12 CTAs per tile with both bands resident, 8 CTAs per tile one band at a
    time; histograms merged over DSMEM) find the SAME exact percentiles as the single-CTA radix select, and numpy's:
    outputs agree to fp32 rounding of the gain"""
    rng = np.random.default_rng(h * 1000 + w)
    T = 3
    bg = np.abs(rng.normal(2.0, 0.7, (T, h, w))).astype(np.float32)
    sig = (bg * 0.8 + rng.normal(0, 0.05, (T, h, w))).astype(np.float32)
    bg[0, : max(1, h // 8), : max(1, w // 8)] = 0
    sig[0, : max(1, h // 8), : max(1, w // 8)] = 0
    sig[1] *= -1.0                                        # negative keys
    bg[2] = 3.25                                          # constant band: max == min
    tb, ts = torch.from_numpy(bg).to(DEV), torch.from_numpy(sig).to(DEV)
    r_clu = features.ratio_2c_match_c_from_sums_outlier(tb, ts).cpu().numpy()
    outs = {}
    for env in ("STARCOP_RATIO_NOCLUSTER", "STARCOP_RATIO_CLUSTER8", "STARCOP_RATIO_CLUSTER12"):
        monkeypatch.setenv(env, "1")
        outs[env] = features.ratio_2c_match_c_from_sums_outlier(tb, ts).cpu().numpy()
        monkeypatch.delenv(env)
    r_one = outs["STARCOP_RATIO_NOCLUSTER"]
    assert np.allclose(r_clu, r_one, rtol=1e-6, atol=1e-6)
    assert np.array_equal(outs["STARCOP_RATIO_CLUSTER8"], r_clu) and np.array_equal(outs["STARCOP_RATIO_CLUSTER12"], r_clu)
    for i in range(T):
        ref = ofeat.ratio_2c_match_c_from_sums_outlier(bg[i].copy(), sig[i].copy())
        assert np.allclose(r_clu[i], ref, rtol=3e-6, atol=3e-6), i


def test_mlr_ratio_vs_reference(golden):
    g = golden("features.npz")
    bands = torch.from_numpy(g["mlr_bands"]).to(DEV)
    tgt = torch.from_numpy(g["mlr_target"]).to(DEV)
    r = features.ratio_MLR_local(bands, tgt).cpu().numpy()
    ref = g["mlr_ratio"]
    assert np.array_equal(r == np.float32(-0.5), ref == np.float32(-0.5))     # nodata / zero-target mask
    assert np.allclose(r, ref, rtol=1e-3, atol=2e-4)
    r5 = features.ratio_MLR_local_5IN(*[b for b in bands], tgt).cpu().numpy()
    assert np.array_equal(r5, r)


def test_emit_rescale_vs_oracle():
    rng = np.random.default_rng(0)
    magic = rng.exponential(120, (70, 100)).astype(np.float32)
    rgb = rng.uniform(0, 50, (3, 70, 100)).astype(np.float32)
    magic[3, 4] = np.nan
    rgb[1, 5, 6] = np.nan
    out = features.emit_rescale(torch.from_numpy(magic).to(DEV), torch.from_numpy(rgb).to(DEV)).cpu().numpy()
    ref = ofeat.emit_rescale(magic, rgb)
    assert out.shape == ref.shape == (4, 64, 96)
    assert np.array_equal(out, ref)
