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
