"""Whole-granule EMIT inference (notebooks/inference_on_raw_EMIT_nc_file.ipynb cells 11-19), device
resident end to end: raw (rows, cols, 285) radiance -> ``mag1c_emit`` (fp64 matched filter in column groups,
mag1c_emit.py:16-90) -> RGB = the bands nearest 640 / 550 / 460 nm -> EMIT rescale to the AVIRIS value range
(emit_tools/emit_dataset.py:62-101) -> ``padded_predict`` (models/utils/padding.py:13-50) of sigmoid(model) --
either the reference's single un-tiled pass or tiles of a given size (BASELINE.json configs[4]'s 256 / 512 / 1024
sweep: every tile is one batch element, so the engine runs a batched forward)."""
import numpy as np
import torch

from . import features, mag1c, tiling


def rgb_band_indices(wavelengths, targets=(640.0, 550.0, 460.0)):
    """index of the band closest to each target wavelength (notebook cell 13)"""
    wl = np.asarray(wavelengths, dtype=np.float64)
    return [int(np.argmin(np.abs(wl - t))) for t in targets]


@torch.no_grad()
def emit_model_input(raw_data, wavelengths, template=None, fwhm=None, fill_value_default=-9999.0, column_step=2, num_iter=30,
                     lut_dir=None):
    """raw (rows, cols, bands) CUDA cube -> ((4, rows32, cols32) model input, mf (rows, cols), albedo)"""
    mf, al = mag1c.mag1c_emit(raw_data, wavelengths, template=template, fwhm=fwhm, fill_value_default=fill_value_default,
                              column_step=column_step, num_iter=num_iter, lut_dir=lut_dir)
    ridx = rgb_band_indices(wavelengths)
    rgb = raw_data[..., torch.as_tensor(ridx, device=raw_data.device)].permute(2, 0, 1).float().contiguous()
    return features.emit_rescale(mf, rgb), mf, al


@torch.no_grad()
def predict_scene(scene, model, tile=None, divisor=32, batch=8, graphed=True):
    """scene: (C, H, W) CUDA model input -> (1, H, W) sigmoid probabilities on the GPU.
    tile=None: ONE reflect-padded pass over the whole scene (``padded_predict``, the notebook's call);
    tile=T: the scene is reflect-padded to a multiple of T (T % 32 == 0) and cut into T x T tiles that run as batch
    elements of `batch` tiles per forward; the central crop is returned.
    graphed: every forward is a replay of a CUDA graph captured per input shape (``ModelModule.forward_graphed``);
    the last, partial batch of tiles is padded to `batch` so that all batches share one graph."""
    model.eval()
    fwd = model.forward_graphed if graphed and hasattr(model, "forward_graphed") else model
    C, H, W = scene.shape
    if tile is None:
        pr, pc = tiling.find_padding(H, divisor), tiling.find_padding(W, divisor)
    else:
        assert tile % divisor == 0, "tile size must be a multiple of 32 (smp check_input_shape)"
        pr, pc = tiling.find_padding(H, tile), tiling.find_padding(W, tile)
    padded = torch.nn.functional.pad(scene[None].float(), (pc[0], pc[1], pr[0], pr[1]), mode="reflect")
    if tile is None:
        prob = torch.sigmoid(fwd(padded))[0]
    else:
        _, _, Hp, Wp = padded.shape
        nh, nw = Hp // tile, Wp // tile
        tiles = padded[0].view(C, nh, tile, nw, tile).permute(1, 3, 0, 2, 4).reshape(nh * nw, C, tile, tile).contiguous()
        nt = nh * nw
        bsz = min(batch, nt)
        outs = []
        for i in range(0, nt, bsz):
            tb = tiles[i:i + bsz]
            if tb.shape[0] < bsz and fwd is not model:          # pad the last batch: one graph for every batch
                tb = torch.cat([tb, tb.new_zeros(bsz - tb.shape[0], C, tile, tile)])
            outs.append(torch.sigmoid(fwd(tb))[:min(bsz, nt - i)])
        prob = torch.cat(outs).view(nh, nw, 1, tile, tile).permute(2, 0, 3, 1, 4).reshape(1, Hp, Wp)
    return prob[:, pr[0]:pr[0] + H, pc[0]:pc[0] + W]
