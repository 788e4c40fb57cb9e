"""Settings tree with the fields the reference reads (scripts/configs/config.yaml;
read sites: model_module.py:26-58, normalizer_module.py:79-112, model_setup.py:7-18).

The reference uses an OmegaConf DictConfig; anything with attribute access and
``in`` works (``"use_weight_loss" not in settings.dataset`` -- model_module.py:46).
"""
from types import SimpleNamespace


class Namespace(SimpleNamespace):
    """SimpleNamespace that also supports ``key in ns`` and ``ns[key]`` like DictConfig."""

    def __contains__(self, k):
        return k in self.__dict__

    def __getitem__(self, k):
        return self.__dict__[k]

    def get(self, k, default=None):
        return self.__dict__.get(k, default)


HYPERSTARCOP_PRODUCTS = ["mag1c", "TOA_AVIRIS_640nm", "TOA_AVIRIS_550nm", "TOA_AVIRIS_460nm"]


def default_settings(input_products=None, pos_weight=1.0, lr=1e-4, use_weight_loss=True, **model_kw):
    """config.yaml defaults; pos_weight=1 is what the paper's HyperSTARCOP runs pass
    (bash/bash_train_example.sh:5), the YAML default is 15."""
    model = dict(train=True, test=False, model_mode="segmentation_output", model_type="unet_semseg",
                 semseg_backbone="mobilenet_v2", num_classes=1, optimizer="adam", lr=lr, lr_decay=0.5,
                 lr_patience=4, loss="BCEWithLogitsLoss", pos_weight=pos_weight, early_stopping_patience=8,
                 model_folder="models")
    model.update(model_kw)
    return Namespace(
        experiment_name="starcop_run", seed=None,
        wandb=Namespace(wandb_project="", wandb_entity="", images_logging="local"),
        dataloader=Namespace(batch_size=32, num_workers=4),
        dataset=Namespace(input_products=list(input_products or HYPERSTARCOP_PRODUCTS),
                          output_products=["labelbinary"], use_weight_loss=use_weight_loss,
                          weight_loss="weight_mag1c", training_size=[128, 128],
                          training_size_overlap=[64, 64], weight_sampling=True),
        model=Namespace(**model),
        training=Namespace(accelerator="gpu", devices=1, max_epochs=15, val_check_interval=.5),
    )


def from_dict(d):
    """nested dict (a parsed config.yaml) -> nested Namespace; lists stay lists"""
    if isinstance(d, dict):
        return Namespace(**{k: from_dict(v) for k, v in d.items()})
    return d


def to_dict(ns):
    if isinstance(ns, SimpleNamespace):
        return {k: to_dict(v) for k, v in ns.__dict__.items()}
    return ns


def load_settings(yaml_path=None, overrides=()):
    """The reference builds its settings with hydra from ``scripts/configs/config.yaml`` plus ``key=value``
    command-line overrides (scripts/train.py:23-26, bash/bash_train_example.sh).  Same result without hydra /
    omegaconf: the YAML (or ``default_settings()`` when no path is given) with dotted overrides applied; values are
    parsed as YAML scalars / lists (``model.pos_weight=1``, ``dataset.input_products=[mag1c,TOA_AVIRIS_640nm]``).
    Like hydra, overriding a key that does not exist is an error unless it is prefixed with ``+``; the ``hydra:``
    block of the file is dropped."""
    import yaml
    if yaml_path is None:
        cfg = to_dict(default_settings())
    else:
        with open(yaml_path) as f:
            cfg = yaml.safe_load(f)
        cfg.pop("hydra", None)
    for ov in overrides:
        if "=" not in ov:
            raise ValueError(f"override {ov!r} is not of the form key=value")
        key, val = ov.split("=", 1)
        add = key.startswith("+")
        key = key.lstrip("+")
        node = cfg
        parts = key.split(".")
        for part in parts[:-1]:
            if part not in node:
                if not add:
                    raise KeyError(f"override {ov!r}: no such config group {part!r} (prefix the key with + to add it)")
                node[part] = {}
            node = node[part]
        if parts[-1] not in node and not add:
            raise KeyError(f"override {ov!r}: no such key (prefix it with + to add it)")
        v = yaml.safe_load(val)
        if isinstance(v, str):                   # PyYAML reads "1e-4" as a string; hydra's grammar reads a float
            try:
                v = float(v) if any(ch in v.lower() for ch in ".e") else int(v)
            except ValueError:
                pass
        node[parts[-1]] = v
    return from_dict(cfg)
