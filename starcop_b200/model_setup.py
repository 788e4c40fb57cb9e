"""Drop-in for ``starcop/model_setup.py:5-20``: same signature, same settings fields."""
import os

from .model_module import ModelModule, load_weights


def get_model(settings, experiment_name=None):
    if settings.model.model_mode == "segmentation_output":
        model = ModelModule(settings)
    elif settings.model.model_mode == "regression_output":
        raise NotImplementedError("the regression module is outside the HyperSTARCOP hot path")
    else:
        raise ValueError(f"unknown model_mode {settings.model.model_mode}")
    if settings.model.test:
        assert experiment_name is not None, "Expermient name must be set on test or deploy mode"
        path_to_models = os.path.join(settings.model.model_folder, experiment_name, "model.pt").replace("\\", "/")
        model.load_state_dict(load_weights(path_to_models))
        print(f"Loaded model weights: {path_to_models}")
    return model
