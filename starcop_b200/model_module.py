"""Drop-in for ``starcop.models.model_module`` (reference file:line cited per method).

``ModelModule`` keeps the reference's LightningModule method surface, constructor signature
(``ModelModule(settings)``), batch-dict contract and state_dict key layout; everything it computes
runs in the hand-written CUDA kernels behind ``include/starcop_b200.h``.  When
``pytorch_lightning`` is importable the class derives from ``pl.LightningModule`` so the
reference's ``Trainer.fit(model, data_module)`` (scripts/train.py:140) accepts it unchanged;
otherwise it is a plain ``torch.nn.Module`` with the same methods.
"""
import os

import numpy as np
import torch

from . import _lib
from . import metrics
from .engine import DT, UNetEngine
from .network import UnetParameters
from .normalizer import DataNormalizer

try:                                                       # pragma: no cover - not in this image
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:                                          # noqa: BLE001
    pl = None
    _Base = torch.nn.Module


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


# --------------------------------------------------------------------------------------------------
# network: parameters in one flat fp32 arena + the CUDA engine
# --------------------------------------------------------------------------------------------------
class _UnetFunction(torch.autograd.Function):
    """logits = network(x): one autograd node for the whole U-Net; backward = the engine's
    hand-written backward pass, gradients delivered as views of the flat gradient arena."""

    @staticmethod
    def forward(ctx, net, x, norm, *params):
        ctx.net = net
        out = net._forward_impl(x, norm, net.training, record=True)
        ctx.generation = net._engine.generation        # the activations this node's backward needs
        return out

    @staticmethod
    def backward(ctx, dlogits):
        net = ctx.net
        net._backward_impl(dlogits, ctx.generation)     # raises if a later recording forward replaced them
        return (None, None, None, *net._grad_views)


class HyperStarcopUnet(UnetParameters):
    """``smp.Unet(mobilenet_v2, in_channels=C, classes=1)`` executed by the sm_100a kernels.

    compute_dtype "f32": fp32 activations + fp32-FMA convolutions (the parity mode: the
    reference trains in fp32, scripts/train.py:120-138).  "bf16": bf16 activations, tcgen05
    tensor-core convolutions with fp32 accumulation (BASELINE.json's throughput mode).
    """

    def __init__(self, in_channels, classes=1, compute_dtype="f32"):
        super().__init__(in_channels, classes)
        self.compute_dtype = compute_dtype
        self._engine = None
        self._flat = None
        self._names = [n for n, _ in self.named_parameters()]

    # ---- flat parameter / gradient arenas ---------------------------------------------------
    def _materialize(self):
        params = list(self.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.StarcopB200Error("starcop_b200 runs on CUDA only: move the module to a GPU (no CPU path)")
        ok = self._flat is not None and self._flat[0].device == dev
        if ok:
            off = 0
            base = self._flat[0].data_ptr()
            for p in params:
                if p.data_ptr() != base + off * 4:
                    ok = False
                    break
                off += (p.numel() + 3) // 4 * 4
        if ok:
            return
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]          # 16 B aligned slices
        total = sum(sizes)
        flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        views, off = [], 0
        for p, s in zip(params, sizes):
            flat_p[off:off + p.numel()].copy_(p.data.reshape(-1).float())
            p.data = flat_p[off:off + p.numel()].view(p.shape)
            views.append(flat_g[off:off + p.numel()].view(p.shape))
            off += s
        self._flat = (flat_p, flat_g)
        self._grad_views = views
        pd = {n: p.data for n, p in self.named_parameters()}
        gd = dict(zip(self._names, views))
        bd = {n: b for n, b in self.named_buffers()}
        self._engine = UNetEngine(pd, bd, gd, self.in_channels, dev, self.compute_dtype)
        self._adam_state = None

    @property
    def flat_params(self):
        self._materialize()
        return self._flat[0]

    @property
    def flat_grads(self):
        self._materialize()
        return self._flat[1]

    def decoder_grad_offset(self):
        """element offset in the flat arenas of the first decoder parameter: [offset:] = decoder + head (the part of
        the gradient the backward pass finishes first)"""
        self._materialize()
        for n, p in self.named_parameters():
            if n.startswith("decoder."):
                return (p.data_ptr() - self._flat[0].data_ptr()) // 4
        return 0

    # ---- forward / backward -------------------------------------------------------------------
    def _forward_impl(self, x, norm, training, record=None):
        """x: (B,C,H,W) tensor, or a producer ``(B, H, W, device, fill)`` whose ``fill(ptr, ld, dtype, stream)`` writes
        the normalised NHWC input straight into the engine's input buffer (the spectral chain's fused pack)."""
        self._materialize()
        eng = self._engine
        producer = None
        if isinstance(x, tuple):
            B, H, W, dev_, producer = x
            C = self.in_channels
            x = None
        else:
            x = x.contiguous().float()
            B, C, H, W = x.shape
            dev_ = x.device
        assert C == self.in_channels, f"expected {self.in_channels} input channels, got {C}"
        if H % 32 or W % 32:                     # smp check_input_shape
            raise RuntimeError(f"Wrong input shape height={H}, width={W}. Expected image height and width "
                               f"divisible by 32.")
        eng.stream = _stream(dev_)
        eng.begin_step(record=training if record is None else record)
        ldin = C if eng.dtype == _lib.SC_F32 else (C + 7) // 8 * 8
        xin = eng.new(B, H, W, C, ld=ldin)
        if producer is not None:
            producer(xin.ptr, ldin, eng.dtype, eng.stream)
        else:
            if norm is None:
                prm, mask = self._identity_norm(dev_, C), 0
            else:
                prm, mask = norm
            _lib.call("sc_normalize_pack", x.data_ptr(), prm[0].data_ptr(), prm[1].data_ptr(), prm[2].data_ptr(),
                      prm[3].data_ptr(), mask, B, C, H, W, xin.ptr, ldin, eng.dtype, 0, eng.stream)
        logits = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev_)
        eng.forward(xin, logits.data_ptr(), training, record)
        if training:
            if getattr(self, "_nbt", None) is None or self._nbt[0].device != dev_:
                self._nbt = [b for n, b in self.named_buffers() if n.endswith("num_batches_tracked")]
            torch._foreach_add_(self._nbt, 1)
        return logits

    def _backward_impl(self, dlogits, generation=None):
        eng = self._engine
        self._flat[1].zero_()
        d = dlogits.contiguous().float()
        eng.stream = _stream(d.device)
        eng.backward(d.data_ptr(), generation)

    def _identity_norm(self, device, C):
        key = (str(device), C)
        if getattr(self, "_idn", None) is None or self._idn[0] != key:
            big = torch.finfo(torch.float32).max
            arr = torch.tensor([[0.] * C, [1.] * C, [-big] * C, [big] * C], dtype=torch.float64, device=device)
            self._idn = (key, arr)
        return self._idn[1]

    def forward(self, x, _norm=None):
        """x: (B,C,H,W) normalised input -> (B,1,H,W) logits (the ``self.network(...)`` call of
        model_module.py:98).  ``_norm`` lets ModelModule fuse normalize_x into the input pack."""
        squeeze = x.dim() == 3
        if squeeze:
            x = x[None]
        params = list(self.parameters())
        # eval-mode forwards never record (the reference allows `model.eval(); model(x)` without no_grad();
        # gradients through running-statistics BatchNorm are not part of the hot path)
        need_grad = self.training and torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if need_grad:
            out = _UnetFunction.apply(self, x, _norm, *params)
        else:
            with torch.no_grad():
                out = self._forward_impl(x, _norm, self.training, record=False)
        return out[0] if squeeze else out

    # ---- roofline probe: the dominant kernel timed alone, live, with CUDA events -----------------
    def bench_dominant_kernel(self, batch, size, iters=20):
        """decoder.blocks.0.conv1 (1376 -> 256 channels, 3x3, at 1/16 resolution): the largest single
        kernel of the step (6.49 GFLOP/tile fprop).  Returns the roofline dict."""
        self._materialize()
        dev = self._flat[0].device
        N, H, W, cin, cout, k = batch, size // 16, size // 16, 1376, 256, 3
        w = dict(self.named_parameters())["decoder.blocks.0.conv1.0.weight"].data
        lib = _lib.load()
        if not lib.sc_tc_supported() or W % 16 or H % 8:
            return None
        cpad = lib.sc_tc_cin_pad(cin)
        wb = torch.empty(cout * k * k * cpad, dtype=torch.bfloat16, device=dev)
        st = _stream(dev)
        _lib.call("sc_tc_pack_weights", w.data_ptr(), wb.data_ptr(), cout, cin, k, k, 0, cpad, cout, st)
        # rotate over several input/output buffers so successive launches do not hit a warm L2
        nbuf = 8
        xs = [torch.randn(N, H, W, cin, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
        ys = [torch.empty(N, H, W, cout, dtype=torch.bfloat16, device=dev) for _ in range(nbuf)]
        def launch(i):
            _lib.call("sc_tc_conv_fprop", xs[i % nbuf].data_ptr(), cin, wb.data_ptr(), ys[i % nbuf].data_ptr(), cout, 0, 0,
                      N, H, W, cin, cout, k, k, 1, 0, st)
        for i in range(3):
            launch(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            launch(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = 2.0 * N * H * W * k * k * cin * cout
        # traffic: dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed
        # `ncu --set full` capture (profiles/r01_b0c1_fprop_ncu_full.json), valid for bs=16 / 512x512 only
        traffic = 51.78e6 if (batch, size) == (16, 512) else None
        return {"bound": "tensor", "kernel": "tc_conv_fprop_kernel<64> (decoder.blocks.0.conv1 fprop)",
                "achieved": flops / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "flop_per_launch": flops,
                "us_per_launch": ms * 1e3, "traffic": traffic, "traffic_unit": "bytes/launch",
                "algorithmic_bytes": 2.0 * N * H * W * (cin + cout) + 2.0 * cout * k * k * cpad}

    # ---- fused optimiser (Adam on the flat arena) -----------------------------------------------
    def adam_step(self, lr, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
        """torch.optim.Adam update of the flat arena.  Step counter and learning rate live on the
        device (sc_adam_step_dev) so the call can sit inside a captured CUDA graph."""
        self._materialize()
        p, g = self._flat
        if self._adam_state is None:
            self._adam_state = {"m": torch.zeros_like(p), "v": torch.zeros_like(p),
                                "step": torch.zeros(1, dtype=torch.int32, device=p.device),
                                "lr": torch.full((1,), float(lr), dtype=torch.float32, device=p.device), "lr_host": float(lr)}
        st = self._adam_state
        if st["lr_host"] != float(lr) and not torch.cuda.is_current_stream_capturing():
            st["lr"].fill_(float(lr))
            st["lr_host"] = float(lr)
        _lib.call("sc_adam_step_dev", p.data_ptr(), g.data_ptr(), st["m"].data_ptr(), st["v"].data_ptr(), p.numel(),
                  st["lr"].data_ptr(), betas[0], betas[1], eps, st["step"].data_ptr(), grad_scale, _stream(p.device))

    def set_lr(self, lr):
        if self._adam_state is not None:
            self._adam_state["lr"].fill_(float(lr))
            self._adam_state["lr_host"] = float(lr)


# --------------------------------------------------------------------------------------------------
# loss: BCEWithLogitsLoss(pos_weight, reduction) + mean(loss * weight_loss), one fused kernel
# --------------------------------------------------------------------------------------------------
def _loss_buffer(B, HW, device):
    """sc_bce_fused's loss accumulator: [0] the fp64 sum, [1] a ticket word, [2:] per-block partials (the kernel
    adds the partials up in block order: deterministic).  Zero-filled, as the C ABI requires of [0] and [1]."""
    return torch.zeros(_lib.load().sc_bce_loss_words(B, HW), dtype=torch.float64, device=device)


class _WeightedBCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, y, w, pos_weight):
        lg = logits.contiguous().float()
        B = lg.shape[0]
        HW = lg.numel() // B
        loss_sum = _loss_buffer(B, HW, lg.device)
        grad = torch.empty_like(lg) if ctx.needs_input_grad[0] else None
        _lib.call("sc_bce_fused", lg.data_ptr(), y.contiguous().float().data_ptr(),
                  w.contiguous().float().data_ptr() if w is not None else 0, float(pos_weight), B, HW,
                  1.0 / lg.numel(), loss_sum.data_ptr(), grad.data_ptr() if grad is not None else 0,
                  0, 0, 0, 0, 0, 0, 0, 0, 0, _stream(lg.device))
        ctx.save_for_backward(grad)
        return (loss_sum[0] / lg.numel()).float()

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return grad * gout, None, None, None


class BCEWithLogitsLoss(torch.nn.Module):
    """``torch.nn.BCEWithLogitsLoss(pos_weight=..., reduction=...)`` of model_module.py:57-58 on the
    fused kernel.  reduction="none" returns the per-pixel loss (used by batch_with_preds :202)."""

    def __init__(self, pos_weight, reduction="none"):
        super().__init__()
        self.pos_weight = pos_weight          # the module's own Parameter -> "loss_function.pos_weight"
        self.reduction = reduction

    def forward(self, logits, target):
        if self.reduction == "mean":
            return _WeightedBCEFunction.apply(logits, target, None, self.pos_weight)
        lg = logits.detach().contiguous().float()
        B = lg.shape[0]
        out = torch.empty_like(lg)
        _lib.call("sc_bce_fused", lg.data_ptr(), target.contiguous().float().data_ptr(), 0, float(self.pos_weight),
                  B, lg.numel() // B, 0.0, 0, 0, 0, 0, 0, 0, 0, out.data_ptr(), 0, 0, 0, _stream(lg.device))
        return out


def weighted_bce(logits, y, weight_loss, pos_weight):
    """torch.mean(BCEWithLogitsLoss(reduction="none")(logits, y) * weight_loss) -- model_module.py:76-79."""
    return _WeightedBCEFunction.apply(logits, y, weight_loss, pos_weight)


class BinaryConfusionMatrix(torch.nn.Module):
    """torchmetrics.ConfusionMatrix(num_classes=2, task="binary") surface used by the reference
    (update / compute / reset, model_module.py:62-63,128,135,149-163): cm[target, pred], int64."""

    def __init__(self):
        super().__init__()
        self.register_buffer("confmat", torch.zeros(2, 2, dtype=torch.long), persistent=False)

    def update(self, preds, target):
        idx = 2 * target.long().flatten() + preds.long().flatten()
        self.confmat += torch.bincount(idx, minlength=4).reshape(2, 2)

    def add_counts(self, counts):
        self.confmat += counts.reshape(2, 2)

    def compute(self):
        return self.confmat.clone()

    def reset(self):
        self.confmat.zero_()

    def forward(self, preds, target):
        before = self.confmat.clone()
        self.update(preds, target)
        return self.confmat - before


def pred_classification(pred_binary):
    """model_module.py:210-212."""
    n_pixels = (10 * np.prod(tuple(pred_binary.shape[-2:]))) / (64 ** 2)
    return (torch.sum(pred_binary, dim=(-1, -2)) > n_pixels).long()


def differences(y_pred_binary, y_gt):
    """model_module.py:268-269."""
    return 2 * y_pred_binary.long() + (y_gt == 1).long()


def configure_architecture(architecture, num_channels, num_classes, extra_settings_model):
    """model_module.py:224-256: the only constructible architecture is smp.Unet(mobilenet_v2)."""
    if architecture == "unet_semseg":
        backbone = extra_settings_model.semseg_backbone
        if backbone != "mobilenet_v2":
            raise Exception(f"No B200 kernel schedule for backbone: {backbone} (HyperSTARCOP uses mobilenet_v2)")
        compute_dtype = extra_settings_model.get("compute_dtype", "f32") if hasattr(extra_settings_model, "get") \
            else getattr(extra_settings_model, "compute_dtype", "f32")
        # the reference loads imagenet encoder weights when num_channels == 3 (model_module.py:242);
        # there is no network here, so a 3-channel model starts from the torchvision initialiser
        return HyperStarcopUnet(num_channels, num_classes, compute_dtype=compute_dtype)
    raise Exception(f"No model implemented for model_type: {architecture}")


def load_weights(path_weights, map_location="cpu"):
    """model_module.py:258-266 (local paths; the reference goes through fsspec for gs://)."""
    if os.path.exists(path_weights):
        return torch.load(path_weights, map_location=map_location)
    raise ValueError(f"Pretrained weights file: {path_weights} does not exists")


class ModelModule(_Base):
    def __init__(self, settings):
        super().__init__()
        if hasattr(self, "save_hyperparameters") and pl is not None:
            self.save_hyperparameters()
        self.settings_model = settings.model
        self.settings_wandb = getattr(settings, "wandb", None)
        self.normalizer = DataNormalizer(settings)
        self.num_classes = self.settings_model.num_classes
        self.num_channels = len(settings.dataset.input_products)
        self.network = configure_architecture(self.settings_model.model_type, self.num_channels,
                                              self.num_classes, self.settings_model)
        self.lr = self.settings_model.lr
        self.lr_decay = self.settings_model.lr_decay
        self.lr_patience = self.settings_model.lr_patience
        self.loss_name = self.settings_model.loss
        use_weight_loss = "use_weight_loss" not in settings.dataset or settings.dataset.use_weight_loss
        if self.settings_model.loss == "BCEWithLogitsLoss":
            self.reduction = "none" if use_weight_loss else "mean"
            self.pos_weight = torch.nn.Parameter(torch.tensor(float(self.settings_model.pos_weight)),
                                                 requires_grad=False)
            self.loss_function = BCEWithLogitsLoss(self.pos_weight, self.reduction)
        else:
            raise NotImplementedError("l1 / mse belong to the regression module (out of the hot path)")
        if self.settings_model.model_mode == "segmentation_output":
            self.confusion_matrix = BinaryConfusionMatrix()
            self.classification_confusion_matrix = BinaryConfusionMatrix()
        elif self.settings_model.model_mode == "regression_output":
            raise NotImplementedError("Not implemented yet")
        self._logged = {}

    # ---- LightningModule surface when pytorch_lightning is absent --------------------------------
    if pl is None:
        @property
        def device(self):
            return next(self.parameters()).device

        def log(self, name, value, *args, **kwargs):
            self._logged[name] = value

        @classmethod
        def load_from_checkpoint(cls, path, settings=None, map_location="cpu", **kw):
            ckpt = torch.load(path, map_location=map_location)
            model = cls(settings)
            model.load_state_dict(ckpt["state_dict"] if "state_dict" in ckpt else ckpt)
            return model
    else:                                                  # pragma: no cover
        def log(self, *args, **kwargs):                    # model_module.py:103-107
            try:
                super().log(*args, **kwargs)
            except Exception as e:                         # noqa: BLE001
                print(f"Bug logging {e}")

    def _pw(self):
        """pos_weight as a Python float without a device sync on the hot path (cached per tensor version)."""
        key = (self.pos_weight.data_ptr(), self.pos_weight._version)
        if getattr(self, "_pw_cache", (None, None))[0] != key:
            self._pw_cache = (key, float(self.pos_weight))
        return self._pw_cache[1]

    # ---- hot path -----------------------------------------------------------------------------------
    def forward(self, x):
        """model_module.py:90-98: network(normalizer.normalize_x(x)); the normalisation is fused
        into the kernel that packs the input to NHWC."""
        if not x.is_cuda:
            raise _lib.StarcopB200Error("starcop_b200 runs on CUDA tensors only (no CPU path)")
        return self.network(x, _norm=self.normalizer.kernel_params(x.device))

    @torch.no_grad()
    def forward_graphed(self, x):
        """Eval-mode ``forward`` replayed from a CUDA graph captured once per input shape (whole-scene / tiled
        inference calls the same shape over and over: ~150 kernel launches become one cudaGraphLaunch).  The graph
        reads the live weights and running statistics (they are re-packed / re-folded inside the graph), so it stays
        valid across optimiser steps.  The returned logits are the graph's static output: consume or clone them before
        the next call with the same shape."""
        assert not self.training, "forward_graphed is an inference path: call .eval() first"
        if not x.is_cuda:
            raise _lib.StarcopB200Error("starcop_b200 runs on CUDA tensors only (no CPU path)")
        graphs = self.__dict__.setdefault("_eval_graphs", {})
        key = (tuple(x.shape), x.dtype, x.device.index)
        ent = graphs.get(key)
        if ent is None:
            static_in = x.clone()
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):
                for _ in range(2):                               # sizes the eval arenas outside the capture
                    self.forward(static_in)
            torch.cuda.current_stream(x.device).wait_stream(side)
            torch.cuda.synchronize(x.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.forward(static_in)
            if len(graphs) >= 8:
                graphs.clear()
            ent = graphs[key] = (g, static_in, out)
        g, static_in, out = ent
        static_in.copy_(x, non_blocking=True)
        g.replay()
        return out

    def training_step(self, batch, batch_idx):
        """model_module.py:69-88."""
        x, y = batch["input"], batch["output"]
        weight_loss = batch["weight_loss"] if self.reduction == "none" else None
        predictions = self.forward(x)
        loss = weighted_bce(predictions, self.normalizer.normalize_y(y), weight_loss, self.pos_weight)
        if (batch_idx % 100) == 0:
            self.log(f"train_{self.loss_name}", loss)
        return loss

    def pred_classification(self, pred_binary):
        return pred_classification(pred_binary)

    def val_step(self, batch, batch_idx, prefix="val"):
        """model_module.py:110-135: loss, pixel confusion matrix (pred = logits >= 0) and tile
        classification confusion matrix from ONE pass over the logits."""
        x, y = batch["input"], batch["output"]
        with torch.no_grad():
            logits = self.forward(x).contiguous()
            y = self.normalizer.normalize_y(y).contiguous().float()
            w = batch["weight_loss"].contiguous().float() if self.reduction == "none" else None
            B = logits.shape[0]
            HW = logits.numel() // B
            dev = logits.device
            loss_sum = _loss_buffer(B, HW, dev)
            cm = torch.zeros(4, dtype=torch.long, device=dev)
            cnt = torch.zeros(B, dtype=torch.long, device=dev)
            _lib.call("sc_bce_fused", logits.data_ptr(), y.data_ptr(), w.data_ptr() if w is not None else 0,
                      self._pw(), B, HW, 0.0, loss_sum.data_ptr(), 0, cm.data_ptr(), cnt.data_ptr(),
                      0, 0, 0, 0, 0, 0, 0, _stream(dev))
            loss = (loss_sum[0] / logits.numel()).float()
            self.log(f"{prefix}_loss", loss, on_epoch=True)
            if self.settings_model.model_mode == "segmentation_output":
                self.confusion_matrix.add_counts(cm)
                n_pixels = (10 * np.prod(tuple(logits.shape[-2:]))) / (64 ** 2)
                pred_cls = (cnt > n_pixels).long()[:, None]
                y_cls = batch["has_plume"][:, None]
                self.classification_confusion_matrix.update(pred_cls, y_cls)
        return loss

    def validation_step(self, batch, batch_idx):
        return self.val_step(batch, batch_idx, prefix="val")

    def test_step(self, batch, batch_idx):
        return self.val_step(batch, batch_idx, prefix="test")

    def val_epoch_end(self, outputs, prefix):
        """model_module.py:147-164."""
        outs = {}
        cm = self.confusion_matrix.compute()
        for fun in metrics.METRICS_CONFUSION_MATRIX:
            self.log(f"{prefix}_{fun.__name__}", fun(cm))
        self.confusion_matrix.reset()
        if self.settings_model.model_mode == "segmentation_output":
            cm = self.classification_confusion_matrix.compute()
            for fun in metrics.METRICS_CONFUSION_MATRIX:
                self.log(f"{prefix}_classification_{fun.__name__}", fun(cm))
            self.classification_confusion_matrix.reset()
        return outs

    def validation_epoch_end(self, outputs):
        self.val_epoch_end(outputs, prefix="val")

    def test_epoch_end(self, outputs):
        self.val_epoch_end(outputs, prefix="test")

    def configure_optimizers(self):
        """model_module.py:172-185."""
        if self.settings_model.optimizer == "adam":
            optimizer = torch.optim.Adam(self.network.parameters(), self.lr)
        else:
            raise Exception(f"No optimizer implemented for : {self.settings_model.optimizer}")
        scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", factor=self.lr_decay,
                                                               patience=self.lr_patience)
        return {"optimizer": optimizer, "lr_scheduler": scheduler, "monitor": "val_loss"}

    def batch_with_preds(self, batch):
        """model_module.py:191-208: every per-pixel product from one fused pass over the logits."""
        with torch.no_grad():
            x = batch["input"]
            B, C, H, W = x.shape
            dev = x.device
            logits = self.forward(x).contiguous()
            batch = batch.copy()
            batch["input_norm"] = self.normalizer.normalize_x(x)
            batch["output_norm"] = self.normalizer.normalize_y(batch["output"])
            y = batch["output_norm"].contiguous().float()
            f = lambda: torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
            li = lambda: torch.empty(B, 1, H, W, dtype=torch.long, device=dev)
            pred, pb, diff = f(), li(), li()
            weighted = self.reduction == "none"
            lpx, lpw = (f(), f()) if weighted else (None, None)
            w = batch["weight_loss"].contiguous().float() if weighted else None
            cnt = torch.zeros(B, dtype=torch.long, device=dev)
            _lib.call("sc_bce_fused", logits.data_ptr(), y.data_ptr(), w.data_ptr() if weighted else 0,
                      self._pw(), B, H * W, 0.0, 0, 0, 0, 0, 0, cnt.data_ptr(), pred.data_ptr(),
                      lpx.data_ptr() if weighted else 0, lpw.data_ptr() if weighted else 0, pb.data_ptr(),
                      diff.data_ptr(), _stream(dev))
            batch["prediction"] = pred
            batch["logits"] = logits
            if weighted:
                batch["loss_per_pixel"] = lpx
                batch["loss_per_pixel_weighted"] = lpw
            batch["pred_binary"] = pb
            batch["differences"] = diff
            n_pixels = (10 * H * W) / (64 ** 2)
            batch["pred_classification"] = (cnt > n_pixels).long()[:, None]
        return batch

    # ---- fused train step (no autograd graph): forward + loss + backward + Adam -------------------
    def train_step_fused(self, batch, lr=None, grad_sync=None):
        """One optimisation step entirely on the engine.  Equivalent to
        ``loss = training_step(batch); loss.backward(); Adam.step()`` of the reference loop.
        ``grad_sync(flat_grads)`` is the data-parallel hook (NCCL all-reduce of the flat arena)."""
        net = self.network
        x, y = batch["input"], batch["output"]
        w = batch["weight_loss"] if self.reduction == "none" else None
        net.train()
        # batch["input"] may be a producer tuple (spectral chain: the pack kernel writes the engine's input buffer)
        logits = net._forward_impl(x, None if isinstance(x, tuple) else self.normalizer.kernel_params(x.device), True)
        B = logits.shape[0]
        n = logits.numel()
        dev = logits.device
        loss_sum = _loss_buffer(B, n // B, dev)
        grad = torch.empty_like(logits)
        _lib.call("sc_bce_fused", logits.data_ptr(), y.contiguous().float().data_ptr(),
                  w.contiguous().float().data_ptr() if w is not None else 0, self._pw(), B, n // B,
                  1.0 / n, loss_sum.data_ptr(), grad.data_ptr(), 0, 0, 0, 0, 0, 0, 0, 0, 0, _stream(dev))
        eng = net._engine
        eng.on_decoder_grads_issued = None
        if grad_sync is not None and hasattr(grad_sync, "early"):
            split = net.decoder_grad_offset()
            eng.on_decoder_grads_issued = lambda main, side: grad_sync.early(net.flat_grads, split, main, side)
        net._backward_impl(grad)
        eng.on_decoder_grads_issued = None
        scale = 1.0
        if grad_sync is not None:
            scale = grad_sync(net.flat_grads)
        net.adam_step(self.lr if lr is None else lr, grad_scale=scale)
        return (loss_sum[0] / n).float()

    def make_graphed_train_step(self, example_batch, grad_sync=None, warmup=2, double_buffer=False):
        """Capture train_step_fused into ONE CUDA graph (every buffer of the step comes from the static
        arenas, the Adam step counter and lr are device resident).  Returns ``step(batch) -> loss``:
        the batch is copied into the graph's static input buffers, the graph is replayed (~600 kernel
        launches, one cudaGraphLaunch), the returned loss tensor is the graph's static output.

        ``double_buffer=True`` captures the step twice, over two sets of input buffers, and adds
        ``step.prefetch(batch)``: the host-to-device copy of the NEXT batch runs on a copy stream while the
        current step computes (what a pinned-memory DataLoader + ``non_blocking`` transfer does in the
        reference loop); ``step()`` without arguments then consumes the oldest prefetched batch."""
        # at least one eager step at THIS batch shape must have run (it sizes the arenas, records the weight-packing
        # plan and caches host state): either `warmup` >= 1 here, or the caller just ran train_step_fused on it
        assert warmup >= 1 or (self.network._engine is not None and self.network._engine._plan_complete), \
            "run one eager train_step_fused on this batch shape first, or pass warmup >= 1"
        dev = example_batch["input"].device
        keys = ["input", "output"] + (["weight_loss"] if self.reduction == "none" else [])
        nset = 2 if double_buffer else 1
        statics = [{k: torch.empty_like(example_batch[k], device=dev) for k in keys} for _ in range(nset)]
        for st in statics:
            for k in keys:
                st[k].copy_(example_batch[k])
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):                      # sizes the arenas, JITs nothing, warms NCCL
                self.train_step_fused(statics[0], grad_sync=grad_sync)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graphs, losses = [], []
        # The step is captured on a HIGH-PRIORITY stream: its kernel nodes (the critical path) then outrank the
        # weight-gradient side stream's nodes whenever both want SMs, and the side work fills what is left.  Measured:
        # 9.50 -> 9.32 ms; the opposite (side stream high) 9.87 ms.  STARCOP_MAIN_PRIO=0 captures on the default stream.
        cap_stream = torch.cuda.Stream(device=dev, priority=-1) if os.environ.get("STARCOP_MAIN_PRIO", "1") != "0" else None
        for st in statics:                               # same arenas, same addresses: only the inputs differ
            g = torch.cuda.CUDAGraph()
            with (torch.cuda.graph(g, stream=cap_stream) if cap_stream is not None else torch.cuda.graph(g)):
                losses.append(self.train_step_fused(st, grad_sync=grad_sync))
            graphs.append(g)
        copy_stream = torch.cuda.Stream(device=dev) if double_buffer else None
        ready = [torch.cuda.Event() for _ in range(nset)]      # H2D of the set finished (copy stream)
        done = [torch.cuda.Event() for _ in range(nset)]       # the step that read the set finished (compute stream)
        state = {"next_fill": 0, "pending": []}

        def prefetch(batch):
            assert double_buffer, "make_graphed_train_step(..., double_buffer=True) is required for prefetch()"
            assert len(state["pending"]) < nset, "both input buffer sets are full: call step() first"
            s_ = state["next_fill"]
            state["next_fill"] = (s_ + 1) % nset
            copy_stream.wait_event(done[s_])             # the previous reader of this set has finished
            with torch.cuda.stream(copy_stream):
                for k in keys:
                    statics[s_][k].copy_(batch[k], non_blocking=True)
                ready[s_].record(copy_stream)
            state["pending"].append(s_)

        def step(batch=None):
            cur = torch.cuda.current_stream(dev)
            if batch is not None:
                assert not state["pending"], "prefetched batches are waiting: call step() without a batch"
                s_ = 0
                for k in keys:
                    statics[0][k].copy_(batch[k], non_blocking=True)
            else:
                assert state["pending"], "nothing prefetched: call step.prefetch(batch) first"
                s_ = state["pending"].pop(0)
                cur.wait_event(ready[s_])
            graphs[s_].replay()
            done[s_].record(cur)
            return losses[s_]
        step.graph, step.static, step.prefetch = graphs[0], statics[0], prefetch
        return step
