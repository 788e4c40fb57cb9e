"""BASELINE.json configs[2]: 125-band AVIRIS-shape BIP cubes -> mag1c matched filter -> HyperSTARCOP U-Net, one
device-resident chain (process_aviris.py:183-219 -> the dataset's products -> model_module.py:98).

The reference materialises GeoTIFF products between the stages (``mag1c.tif``, ``TOA_AVIRIS_{640,550,460}nm.tif``,
``weight_mag1c.tif``) and re-reads them in the DataLoader.  Here the matched filter runs straight on the cube's
memory (groups = detector columns) and ONE pack kernel (``sc_chain_pack``) turns its output plus the three RGB bands
into the network's normalised NHWC input inside the engine's own input buffer -- the (B,4,H,W) "input" batch of the
DataLoader contract is never written unless asked for."""
import numpy as np
import torch

from . import _lib, mag1c


def rgb_bands(band_centers_nm):
    """indices of the cube bands behind TOA_AVIRIS_640nm / 550nm / 460nm (nearest centre)"""
    c = np.asarray(band_centers_nm, dtype=np.float64)
    return tuple(int(np.argmin(np.abs(c - t))) for t in (640.0, 550.0, 460.0))


@torch.no_grad()
def cube_batch(model, cube, template, band_slice, rgb_idx, output=None, num_iter=30, materialize_input=False):
    """cube: (B,H,W,C) CUDA f32 BIP -> a batch dict for ``ModelModule.train_step_fused`` / ``forward``: "input" is a
    producer that fills the engine's NHWC buffer (or the raw (B,4,H,W) tensor when ``materialize_input``),
    "weight_loss" = weight_mag1c (B,1,H,W), "output" = the given labels, "mag1c" = the filter output (B,H,W)."""
    if not cube.is_cuda:
        raise _lib.StarcopB200Error("starcop_b200 runs on CUDA tensors only (no CPU path)")
    cube = cube.contiguous()
    B, H, W, C = cube.shape
    dev = cube.device
    mf, _ = mag1c.mag1c_tiles(cube, template, band_slice, num_iter=num_iter)
    prm, mask = model.normalizer.kernel_params(dev)
    assert mask == 0, "sc_chain_pack implements the integer-factor normaliser of the HyperSTARCOP products"
    weight = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
    raw = torch.empty(B, 4, H, W, dtype=torch.float32, device=dev) if materialize_input else None
    st = torch.cuda.current_stream(dev).cuda_stream
    r, g, b = rgb_idx

    def fill(ptr, ld, dtype, stream):
        _lib.call("sc_chain_pack", mf.data_ptr(), cube.data_ptr(), C, r, g, b, prm[0].data_ptr(), prm[1].data_ptr(),
                  prm[2].data_ptr(), prm[3].data_ptr(), B, H * W, ptr, ld, dtype, 0, 0, stream)

    # weight_mag1c (and the raw batch if requested) in one launch; the NHWC pack happens inside the forward
    _lib.call("sc_chain_pack", mf.data_ptr(), cube.data_ptr(), C, r, g, b, prm[0].data_ptr(), prm[1].data_ptr(), prm[2].data_ptr(),
              prm[3].data_ptr(), B, H * W, 0, 8, _lib.SC_F32, raw.data_ptr() if raw is not None else 0, weight.data_ptr(), st)
    out = {"input": raw if materialize_input else (B, H, W, dev, fill), "weight_loss": weight, "mag1c": mf}
    if output is not None:
        out["output"] = output
    return out
