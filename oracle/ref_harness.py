"""Import harness for the UNMODIFIED reference (facebookresearch/fast3r @ /root/reference).

TEST / BENCH INFRASTRUCTURE ONLY. This file is used in the build container to (a) validate the
oracle restatement in ``oracle/fast3r_oracle.py`` and (b) generate the golden fixtures under
``tests/golden/``; and by ``bench.py --impl reference`` / the library-bar leg to run the reference itself.
``/root/reference`` does not exist on the GPU box: there the harness falls back to ``oracle/_ref`` (a verbatim,
git-ignored copy of the hot-path modules made by ``oracle/make_ref.py``).  The product never imports this module.

Two in-memory stubs are needed because ``fast3r/models/fast3r.py:13`` imports omegaconf and
``fast3r/utils/__init__.py:7-11`` pulls hydra/lightning.
No reference file is modified or copied.
"""
import logging
import os
import sys
import types

def _default_root() -> str:
    if os.path.isdir("/root/reference/fast3r"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = os.environ.get("FAST3R_REFERENCE_ROOT") or _default_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "fast3r"))


def import_reference():
    """Returns (Fast3R, inference) from the real reference package."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    if "omegaconf" not in sys.modules:
        om = types.ModuleType("omegaconf")

        class DictConfig(dict):
            pass

        class OmegaConf:
            to_container = staticmethod(lambda x, **k: dict(x))

        om.DictConfig, om.OmegaConf = DictConfig, OmegaConf
        sys.modules["omegaconf"] = om
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import fast3r  # noqa: F401  (the real package)

    if "fast3r.utils" not in sys.modules or not hasattr(sys.modules["fast3r.utils"], "pylogger"):
        fu = types.ModuleType("fast3r.utils")
        fu.__path__ = [os.path.join(REFERENCE_ROOT, "fast3r", "utils")]  # real submodules, but not the hydra-laden __init__
        pl = types.ModuleType("fast3r.utils.pylogger")

        class RankedLogger(logging.LoggerAdapter):
            def __init__(self, name=__name__, rank_zero_only=False, extra=None):
                super().__init__(logging.getLogger(name), extra or {})

        pl.RankedLogger = RankedLogger
        fu.pylogger = pl
        sys.modules["fast3r.utils"] = fu
        sys.modules["fast3r.utils.pylogger"] = pl
    from fast3r.models.fast3r import Fast3R
    from fast3r.dust3r.inference_multiview import inference

    return Fast3R, inference


def import_reference_lit_module(roma_registration=None):
    """The reference's Lightning module (fast3r/models/multiview_dust3r_module.py) with in-memory stand-ins for the
    training-stack packages this image lacks (module lines 1-29: roma, lightning, torchmetrics, pl_bolts, open3d).
    `roma_registration(x, y, compute_scaling=True) -> (R, t, s)` becomes ``roma.rigid_points_registration`` - roma is
    not installable here, see oracle/geometry_oracle.py."""
    import torch
    import torch.nn as nn

    import_reference()

    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        @property
        def device(self):
            return torch.device("cpu")

    class _Metric(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def update(self, *a, **k):
            pass

        def compute(self):
            return torch.tensor(0.0)

    class BaseAggregator(_Metric):
        def __init__(self, fn=None, default_value=None, nan_strategy=None, state_name="value", **k):
            super().__init__()
            setattr(self, state_name, default_value)

    for name in ("open3d", "pl_bolts", "pl_bolts.optimizers", "lightning.pytorch", "lightning.pytorch.loggers"):
        mod(name)
    roma = mod("roma")
    if roma_registration is not None:
        roma.rigid_points_registration = roma_registration
    mod("lightning", LightningModule=LightningModule)
    mod("lightning.pytorch.loggers.wandb", WandbLogger=type("WandbLogger", (), {}))
    mod("torchmetrics", MaxMetric=_Metric, MeanMetric=_Metric, MinMetric=_Metric, SumMetric=_Metric, Metric=_Metric)
    mod("torchmetrics.aggregation", BaseAggregator=BaseAggregator)
    mod("pl_bolts.optimizers.lr_scheduler", LinearWarmupCosineAnnealingLR=type("LinearWarmupCosineAnnealingLR", (), {}))
    import fast3r.models.multiview_dust3r_module as lit_mod

    return lit_mod
