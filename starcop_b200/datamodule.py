"""Feeder with the method surface of ``starcop.data.datamodule.Permian2019DataModule`` (datamodule.py:68-306) and
the batch contract of ``STARCOPDataset.__getitem__`` (dataset.py:40-102) over SYNTHETIC scenes (there is no dataset
and no GDAL here; the GeoTIFF readers are out of scope).  What is reproduced bit-exactly is the
host-side index arithmetic around the reader:

* 512 x 512 scenes are cut into training chips with ``tiled_dataframe`` semantics (datamodule.py:17-64 on
  georeader ``create_windows``: ``tiling.tiled_records``), ids ``{id}_r{row}_c{col}_w{w}_h{h}``,
  ``has_plume = frac_positives > 10 / 64**2``;
* ``train_dataloader`` draws chips with a ``WeightedRandomSampler`` over ``add_sample_weight`` (:273-292, :309-315),
  ``val`` / ``test`` loaders iterate the full 512 x 512 scenes in order (:226-229, :294-306);
* every item is the dict {"input" (C,H,W) f32 raw products, "output" (1,H,W) f32 in {0,1}, "weight_loss" (1,H,W),
  "id" str, "has_plume" int}; ``default_collate`` batches it (pinned memory: the trainer copies batches to the GPU
  and augments them THERE, ``augment.TrainAugmentation``, instead of in CPU workers)."""
import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, WeightedRandomSampler

from . import synthetic, tiling


class SyntheticSceneDataset(Dataset):
    """map-style dataset over (scene, window) records; ``window`` None = the whole scene"""

    def __init__(self, scenes, records, use_weight_loss=True):
        self.scenes = scenes                    # list of dicts of full-scene tensors
        self.records = records                  # list of dicts: scene index, id, window, has_plume
        self.use_weight_loss = use_weight_loss

    def __len__(self):
        return len(self.records)

    @property
    def dataframe(self):                        # the reference exposes a pandas frame; the fields used downstream
        return self.records

    def __getitem__(self, idx):
        rec = self.records[idx]
        sc = self.scenes[rec["scene"]]
        if rec.get("window") is None:
            sl = (slice(None), slice(None), slice(None))
        else:
            r, c, h, w = rec["window"]
            sl = (slice(None), slice(r, r + h), slice(c, c + w))
        out = {"input": sc["input"][sl].float(), "output": sc["output"][sl].float()}
        if self.use_weight_loss:
            out["weight_loss"] = sc["weight_loss"][sl].float()
        out["id"] = str(rec["id"])
        out["has_plume"] = int(rec["has_plume"])
        return out


class SyntheticDataModule:
    """``Permian2019DataModule`` surface: ``prepare_data()``, ``train_dataloader()``, ``val_dataloader()``,
    ``test_dataloader()``, ``train_dataset`` / ``val_dataset`` / ``test_dataset``, ``batch_size``, ``num_workers``."""

    def __init__(self, settings, n_train_scenes=9, n_test_scenes=4, scene_size=512, seed=0):
        ds = settings.dataset
        self.settings = settings
        self.batch_size = settings.dataloader.batch_size
        self.num_workers = settings.dataloader.num_workers
        self.input_products = list(ds.input_products)
        self.output_products = list(ds.output_products)
        self.training_size = tuple(ds.training_size)
        self.training_size_overlap = tuple(ds.training_size_overlap)
        self.weight_sampling = bool(ds.get("weight_sampling", True)) if hasattr(ds, "get") else bool(getattr(ds, "weight_sampling", True))
        self.use_weight_loss = ("use_weight_loss" not in ds) or bool(ds.use_weight_loss)
        self.n_train_scenes, self.n_test_scenes, self.scene_size, self.seed = n_train_scenes, n_test_scenes, scene_size, seed
        self.train_dataset = self.val_dataset = self.test_dataset = self.train_dataset_non_tiled = None

    def _scenes(self, n, seed):
        b = synthetic.hyperstarcop_batch(n, size=self.scene_size, seed=seed, channels=len(self.input_products))
        return [{"input": b["input"][i], "output": b["output"][i], "weight_loss": b["weight_loss"][i],
                 "id": b["id"][i], "has_plume": int(b["has_plume"][i])} for i in range(n)]

    def prepare_data(self):
        S = self.scene_size
        train_scenes = self._scenes(self.n_train_scenes, self.seed + 1)
        test_scenes = self._scenes(self.n_test_scenes, self.seed + 2)
        base = [{"id": s["id"], "scene": i} for i, s in enumerate(train_scenes)]
        tiled = tiling.tiled_records(base, [s["output"] for s in train_scenes], self.training_size,
                                     self.training_size_overlap, scene_shape=(S, S))
        self.train_dataset = SyntheticSceneDataset(train_scenes, tiled, self.use_weight_loss)
        whole = lambda scenes: [{"id": s["id"], "scene": i, "window": None, "has_plume": s["has_plume"]} for i, s in enumerate(scenes)]
        self.train_dataset_non_tiled = SyntheticSceneDataset(train_scenes, whole(train_scenes), self.use_weight_loss)
        self.val_dataset = SyntheticSceneDataset(test_scenes, whole(test_scenes), self.use_weight_loss)   # the reference validates on the test csv
        self.test_dataset = self.val_dataset

    def _loader(self, dataset, batch_size, sampler=None, shuffle=False, num_workers=None):
        return DataLoader(dataset, batch_size=batch_size, sampler=sampler, shuffle=shuffle,
                          num_workers=self.num_workers if num_workers is None else num_workers,
                          pin_memory=torch.cuda.is_available())

    def train_dataloader(self, num_workers=None, batch_size=None, generator=None):
        batch_size = batch_size or self.batch_size
        if self.weight_sampling:                                           # datamodule.py:278-286
            w = tiling.add_sample_weight([r["has_plume"] for r in self.train_dataset.records])
            sampler = WeightedRandomSampler(torch.as_tensor(np.asarray(w, dtype=np.float64)), num_samples=len(self.train_dataset),
                                            replacement=True, generator=generator)
            return self._loader(self.train_dataset, batch_size, sampler=sampler, num_workers=num_workers)
        return self._loader(self.train_dataset, batch_size, shuffle=True, num_workers=num_workers)

    def val_dataloader(self, num_workers=None, batch_size=None):
        return self._loader(self.val_dataset, batch_size or self.batch_size, num_workers=num_workers)

    def test_dataloader(self, num_workers=None, batch_size=None):
        return self._loader(self.test_dataset, batch_size or self.batch_size, num_workers=num_workers)

    def test_plot_dataloader(self, batch_size, num_workers=0):
        return self._loader(self.test_dataset, batch_size, num_workers=num_workers)


def get_dataset(settings, **kw):
    """``starcop.dataset_setup.get_dataset(settings)`` (dataset_setup.py:3) for the synthetic feeder."""
    return SyntheticDataModule(settings, **kw)
