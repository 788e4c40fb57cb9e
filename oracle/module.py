"""Restatement of ``starcop/models/model_module.py`` ``ModelModule`` (a
``pl.LightningModule``; pytorch_lightning is absent here so it is a plain
``torch.nn.Module`` with the same method surface) and ``starcop/model_setup.py:5-20``.

Test infrastructure only (see ``oracle/__init__.py``).
"""
import torch

from . import loss_metrics as lm
from .normalizer import normalize_x, normalize_y, normalizer_params
from .unet import Unet


class OracleModelModule(torch.nn.Module):
    def __init__(self, settings):
        super().__init__()
        sm = settings.model
        self.settings_model = sm
        self.input_products = list(settings.dataset.input_products)
        self.output_products = list(settings.dataset.output_products)
        self.num_channels = len(self.input_products)
        assert sm.model_type == "unet_semseg" and sm.semseg_backbone == "mobilenet_v2"
        self.network = Unet(in_channels=self.num_channels, classes=sm.num_classes)   # :238-251
        self.lr, self.lr_decay, self.lr_patience = sm.lr, sm.lr_decay, sm.lr_patience
        ds = settings.dataset
        use_weight_loss = (not hasattr(ds, "use_weight_loss")) or ds.use_weight_loss  # :46
        assert sm.loss == "BCEWithLogitsLoss"
        self.reduction = "none" if use_weight_loss else "mean"
        self.pos_weight = torch.nn.Parameter(torch.tensor(float(sm.pos_weight)), requires_grad=False)
        self.cm = torch.zeros(2, 2, dtype=torch.long)
        self.cm_cls = torch.zeros(2, 2, dtype=torch.long)

    def forward(self, x):                                                      # :90-98
        return self.network(normalize_x(x, self.input_products))

    def _loss(self, logits, y, batch):
        l = lm.bce_with_logits_elementwise(logits, y, self.pos_weight)
        return torch.mean(l * batch["weight_loss"]) if self.reduction == "none" else torch.mean(l)

    def training_step(self, batch, batch_idx=0):                               # :69-88
        return self._loss(self.forward(batch["input"]), normalize_y(batch["output"], self.output_products), batch)

    def val_step(self, batch, batch_idx=0):                                    # :110-135
        logits = self.forward(batch["input"])
        y = normalize_y(batch["output"], self.output_products)
        loss = self._loss(logits, y, batch)
        pred = lm.pred_val(logits)
        self.cm += lm.confusion_matrix(pred, y.long())
        self.cm_cls += lm.confusion_matrix(lm.pred_classification(pred), batch["has_plume"][:, None])
        return loss

    def val_epoch_end(self, prefix="val"):                                     # :147-164
        out = {f"{prefix}_{f.__name__}": f(self.cm) for f in lm.METRICS_CONFUSION_MATRIX}
        out.update({f"{prefix}_classification_{f.__name__}": f(self.cm_cls) for f in lm.METRICS_CONFUSION_MATRIX})
        self.cm.zero_(); self.cm_cls.zero_()
        return out

    def configure_optimizers(self):                                            # :172-185
        opt = torch.optim.Adam(self.network.parameters(), self.lr)
        sch = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=self.lr_decay,
                                                         patience=self.lr_patience)
        return {"optimizer": opt, "lr_scheduler": sch, "monitor": "val_loss"}

    def batch_with_preds(self, batch):                                         # :191-208
        logits = self(batch["input"])
        pred = torch.sigmoid(logits)
        batch = dict(batch)
        batch["input_norm"] = normalize_x(batch["input"], self.input_products)
        batch["output_norm"] = normalize_y(batch["output"], self.output_products)
        batch["prediction"], batch["logits"] = pred, logits
        if self.reduction == "none":
            batch["loss_per_pixel"] = lm.bce_with_logits_elementwise(logits, batch["output_norm"], self.pos_weight)
            batch["loss_per_pixel_weighted"] = batch["weight_loss"] * batch["loss_per_pixel"]
        batch["pred_binary"] = (pred > .5).long()
        batch["differences"] = lm.differences(batch["pred_binary"], batch["output_norm"].long())
        batch["pred_classification"] = lm.pred_classification(batch["pred_binary"])
        return batch


def get_model(settings, experiment_name=None):
    """model_setup.py:5-20 (segmentation mode, no checkpoint loading in the oracle)."""
    assert settings.model.model_mode == "segmentation_output"
    return OracleModelModule(settings)
