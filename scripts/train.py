#!/usr/bin/env python
"""``python -m scripts.train key=value ...`` -- the reference's entry point (scripts/train.py:23-165) without hydra,
pytorch-lightning, W&B or gs://: reads the same YAML (``scripts/configs/config.yaml`` of the reference when
``--config`` points at it, else the built-in defaults), applies the same dotted ``key=value`` overrides
(bash/bash_train_example.sh), builds the data module and the B200 ``ModelModule`` and runs the fit loop of
``starcop_b200.trainer``.  There is no dataset in this environment: the data module is the synthetic feeder with the
``Permian2019DataModule`` surface (``starcop_b200.datamodule``).

    python -m scripts.train model.pos_weight=1 training.max_epochs=2 dataloader.batch_size=16 +model.compute_dtype=bf16
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=None, help="path to a config.yaml (the reference's scripts/configs/config.yaml)")
    ap.add_argument("--experiment-dir", default=None)
    ap.add_argument("--train-scenes", type=int, default=9)
    ap.add_argument("--test-scenes", type=int, default=4)
    ap.add_argument("overrides", nargs="*")
    a = ap.parse_args(argv)
    from starcop_b200 import settings as S
    from starcop_b200 import trainer
    from starcop_b200.datamodule import get_dataset
    st = S.load_settings(a.config, a.overrides)
    exp = a.experiment_dir or os.path.join("experiments", str(st.experiment_name), time.strftime("%Y-%m-%d_%H-%M"))
    dm = get_dataset(st, n_train_scenes=a.train_scenes, n_test_scenes=a.test_scenes)
    model, tr, report = trainer.train(st, data_module=dm, experiment_path=exp)
    print({k: v for k, v in report.items() if k in ("iou", "f1score", "precision", "recall", "FPR_no_plume")})
    print("checkpoints:", tr.best_path, os.path.join(exp, "final_checkpoint_model.ckpt"))


if __name__ == "__main__":
    main()
