"""CPU oracle for the STARCOP hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``starcop_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / the timed CPU reference.

Every function here restates, in plain torch-CPU / numpy, the arithmetic of one
function on the reference's hot path and cites the reference
file:line it follows.  The reference is pure Python, so the restatement is Python.

Pinning: the reference holds no tests or golden vectors.  The
oracle is therefore pinned against the reference ITSELF, run in the build container:
``tests/golden/make_golden.py`` imports the reference's own modules from
``/root/reference`` (``starcop.metrics``, ``starcop.data.normalizer_module``,
``starcop.data.feature_extration``, ``starcop.models.mag1c`` and
``starcop.models.model_module`` under stub ``pytorch_lightning`` / ``torchmetrics`` /
``segmentation_models_pytorch`` modules) and writes their outputs on seeded inputs to
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every oracle function
against those files.  Third-party arithmetic that is not under ``/root/reference``
(segmentation_models_pytorch, torchmetrics 0.10, kornia 0.6.7, georeader) is
restated from its published behaviour and pinned by: the parameter count the
reference's notebook prints (6 629 233 + 17 frozen = 26.517 MB), the real
``torchvision.models.MobileNetV2`` (present in this image) for the encoder, and the
441 = 9 x 49 window count the notebook prints for ``create_windows``.
"""
