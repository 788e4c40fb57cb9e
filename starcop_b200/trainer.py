"""PL-less fit loop with the step order of ``Trainer.fit(model, data_module)`` as the reference configures it
(scripts/train.py:87-143):

* epochs of ``data_module.train_dataloader()`` batches -> ``training_step`` -> backward -> Adam (here the fused
  device step ``ModelModule.train_step_fused``: forward + weighted BCE + hand-written backward + Adam in one pass,
  optionally captured as ONE CUDA graph) with the training augmentation applied on the GPU
  (``augment.TrainAugmentation``, datamodule.py:128-134);
* validation every ``training.val_check_interval`` (a fraction of an epoch or a batch count, PL semantics) over
  ``val_dataloader()``: ``validation_step`` per batch, ``validation_epoch_end`` -> the logged metric set
  (model_module.py:147-164); ``val_loss`` is the batch-size-weighted epoch mean (``self.log(..., on_epoch=True)``);
* ``ReduceLROnPlateau(mode="min", factor=lr_decay, patience=lr_patience)`` stepped on ``val_loss``
  (model_module.py:178-185) and pushed into the device-resident learning rate (``set_lr``: the captured graph
  picks it up without re-capture);
* ``ModelCheckpoint(monitor="val_loss", mode="min", save_top_k=1)`` (train.py:90-96) + the final checkpoint
  (:143) in Lightning's file layout (``{"state_dict", "epoch", "global_step", ...}``: the state_dict keys are the
  reference's, so ``ModelModule.load_from_checkpoint`` of either implementation reads it); ``resume`` restores
  weights, Adam moments, step counter, scheduler and best score.
EarlyStopping is constructed but NOT registered in the reference (train.py:98-114), so there is none here."""
import json
import os
import time

import torch

from .augment import TrainAugmentation


def to_device(batch, device):
    """starcop/torch_utils.py:to_device with non_blocking copies (the loaders pin their batches)"""
    return {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}


class Trainer:
    def __init__(self, max_epochs=15, val_check_interval=0.5, checkpoint_dir=None, log_every_n_steps=10, augment=True,
                 seed=0, use_cuda_graph=True, grad_sync=None, log=print):
        self.max_epochs, self.val_check_interval = max_epochs, val_check_interval
        self.checkpoint_dir, self.log_every_n_steps = checkpoint_dir, log_every_n_steps
        self.augmentation = TrainAugmentation(seed=seed) if augment is True else (augment or None)
        self.use_cuda_graph, self.grad_sync, self.log = use_cuda_graph, grad_sync, log
        self.global_step, self.current_epoch = 0, 0
        self.best_score, self.best_path = None, None
        self.history = []                                       # one dict per validation run
        self._graphed = {}                                      # batch shape -> graphed step

    # ---- validation --------------------------------------------------------------------------------
    @torch.no_grad()
    def validate(self, model, loader):
        model.eval()
        dev = model.device
        tot, n = 0.0, 0
        for i, batch in enumerate(loader):
            b = to_device(batch, dev)
            loss = model.validation_step(b, i)
            bs = b["input"].shape[0]
            tot += float(loss) * bs
            n += bs
        model.validation_epoch_end(None)
        logged = {k: (float(v) if torch.is_tensor(v) else v) for k, v in model._logged.items()} if hasattr(model, "_logged") else {}
        logged["val_loss"] = tot / max(n, 1)
        model.train()
        return logged

    # ---- checkpoints -------------------------------------------------------------------------------
    def _ckpt(self, model, scheduler):
        net = model.network
        st = net._adam_state
        return {"state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
                "epoch": self.current_epoch, "global_step": self.global_step,
                "pytorch-lightning_version": "starcop_b200 (Lightning checkpoint layout)",
                "optimizer_states": [{"flat_exp_avg": st["m"].cpu(), "flat_exp_avg_sq": st["v"].cpu(),
                                      "step": int(st["step"].item()), "lr": float(st["lr"].item())}] if st else [],
                "lr_schedulers": [scheduler.state_dict()],
                "callbacks": {"ModelCheckpoint": {"best_model_score": self.best_score, "best_model_path": self.best_path}}}

    def save_checkpoint(self, model, path, scheduler):
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        torch.save(self._ckpt(model, scheduler), path)

    def _resume(self, model, scheduler, path):
        ck = torch.load(path, map_location="cpu", weights_only=False)
        model.load_state_dict(ck["state_dict"])
        self.current_epoch, self.global_step = ck["epoch"], ck["global_step"]
        cb = ck.get("callbacks", {}).get("ModelCheckpoint", {})
        self.best_score, self.best_path = cb.get("best_model_score"), cb.get("best_model_path")
        if ck.get("lr_schedulers"):
            scheduler.load_state_dict(ck["lr_schedulers"][0])
        if ck.get("optimizer_states"):
            o = ck["optimizer_states"][0]
            net = model.network
            net.adam_step(0.0, grad_scale=0.0)                  # allocates the state (lr 0: parameters untouched)
            st = net._adam_state
            st["m"].copy_(o["flat_exp_avg"]); st["v"].copy_(o["flat_exp_avg_sq"])
            st["step"].fill_(o["step"])
            net.set_lr(o["lr"])

    # ---- fit ---------------------------------------------------------------------------------------
    def fit(self, model, data_module, resume_from_checkpoint=None):
        dev = model.device
        cfg = model.configure_optimizers()
        scheduler, host_opt = cfg["lr_scheduler"], cfg["optimizer"]
        model.train()
        if resume_from_checkpoint:
            self._resume(model, scheduler, resume_from_checkpoint)
        start_epoch = self.current_epoch
        for epoch in range(start_epoch, self.max_epochs):
            self.current_epoch = epoch
            loader = data_module.train_dataloader()
            nb = len(loader)
            vci = self.val_check_interval
            every = max(1, int(nb * vci)) if isinstance(vci, float) and vci <= 1.0 else int(vci)
            t0 = time.time()
            for i, batch in enumerate(loader):
                b = to_device(batch, dev)
                if self.augmentation is not None:
                    b = self.augmentation(b)
                b = {k: v for k, v in b.items() if not k.startswith("_")}
                if self.use_cuda_graph:
                    key = tuple(b["input"].shape)
                    if key not in self._graphed:
                        # make_graphed_train_step runs ONE eager optimisation step on this batch (it sizes the
                        # arenas) before capturing: that step is this batch's step
                        loss = model.train_step_fused(b, grad_sync=self.grad_sync)
                        self._graphed[key] = model.make_graphed_train_step(b, grad_sync=self.grad_sync, warmup=0)
                    else:
                        loss = self._graphed[key](b)
                else:
                    loss = model.train_step_fused(b, grad_sync=self.grad_sync)
                self.global_step += 1
                if self.global_step % self.log_every_n_steps == 0:
                    self.log(f"epoch {epoch} step {self.global_step} train_{model.loss_name} {float(loss):.5f}")
                if (i + 1) % every == 0 or (i + 1) == nb and every > nb:
                    logged = self.validate(model, data_module.val_dataloader())
                    logged.update(epoch=epoch, global_step=self.global_step)
                    self.history.append(logged)
                    scheduler.step(logged["val_loss"])
                    model.network.set_lr(host_opt.param_groups[0]["lr"])
                    self.log(f"epoch {epoch} step {self.global_step} val_loss {logged['val_loss']:.5f} "
                             f"lr {host_opt.param_groups[0]['lr']:.2e} val_iou {logged.get('val_iou', float('nan')):.4f}")
                    if self.checkpoint_dir and (self.best_score is None or logged["val_loss"] < self.best_score):
                        if self.best_path and os.path.exists(self.best_path):
                            os.remove(self.best_path)               # save_top_k = 1
                        self.best_score = logged["val_loss"]
                        self.best_path = os.path.join(self.checkpoint_dir, f"epoch={epoch}-step={self.global_step}.ckpt")
                        self.save_checkpoint(model, self.best_path, scheduler)
            self.log(f"epoch {epoch}: {nb} batches in {time.time() - t0:.1f} s")
        self.current_epoch = self.max_epochs
        self._scheduler = scheduler
        return self.history


def train(settings, data_module=None, experiment_path=".", log=print, **trainer_kw):
    """``scripts/train.py:train`` without hydra / W&B / gs:// (train.py:23-165): dataset, model, checkpointing,
    fit, final checkpoint, validation report of the test scenes.  Returns (model, trainer, report)."""
    from . import validation
    from .datamodule import get_dataset
    from .model_setup import get_model
    os.makedirs(experiment_path, exist_ok=True)
    ckpt_dir = os.path.join(experiment_path, "checkpoint")
    os.makedirs(ckpt_dir, exist_ok=True)
    seed = None if settings.seed in (None, "None") else int(settings.seed)
    if seed is not None:
        torch.manual_seed(seed)
    dm = data_module if data_module is not None else get_dataset(settings)
    dm.prepare_data()
    settings.model.test, settings.model.train = False, True
    model = get_model(settings, settings.experiment_name).to("cuda")
    tr = Trainer(max_epochs=settings.training.max_epochs, val_check_interval=settings.training.val_check_interval,
                 checkpoint_dir=ckpt_dir, log_every_n_steps=getattr(settings.training, "train_log_every_n_steps", 10),
                 seed=seed or 0, log=log, **trainer_kw)
    resume = None
    if getattr(settings, "resume_from_checkpoint", False):
        cands = sorted(f for f in os.listdir(ckpt_dir) if f.endswith(".ckpt"))
        resume = os.path.join(ckpt_dir, cands[-1]) if cands else None
    tr.fit(model, dm, resume_from_checkpoint=resume)
    tr.save_checkpoint(model, os.path.join(experiment_path, "final_checkpoint_model.ckpt"), tr._scheduler)
    rows, gcm, sweep = validation.run_validation(model, dm.test_plot_dataloader(batch_size=1))
    rows, report = validation.aggregate(rows, gcm, sweep)
    with open(os.path.join(experiment_path, "results_agg.json"), "w") as fh:
        json.dump({k: (v.tolist() if torch.is_tensor(v) else v) for k, v in report.items() if k != "thresholded"}, fh)
    return model, tr, report
