"""ddnerf_b200: the DDNeRF / mip-NeRF per-ray hot path on B200."""
import importlib
import sys

_MODEL_MODULES = ("models", "samplers", "dd_utils", "base_architectures", "helpers")
_UTIL_MODULES = ("math_utils", "nerf_helpers", "volume_rendering_utils")


def install_as_reference():
    """Make ``from models import models`` / ``from general_utils import ...`` -- the imports of the
    reference's unchanged drivers (train_model.py:4,14-15, eval_nerf.py:6,10, render_video.py:10,13) --
    resolve to this package's CUDA-backed mirrors.

    If the reference's own ``general_utils`` package is importable (its tree is on ``sys.path``) it is
    kept for everything outside the hot path (``cfgnode``, ``learning_rate_decay`` ...) and only the
    three hot modules are replaced inside it; otherwise this package's ``general_utils`` stands in.
    """
    pkg = importlib.import_module("ddnerf_b200.models")
    sys.modules["models"] = pkg
    for name in _MODEL_MODULES:
        sys.modules[f"models.{name}"] = importlib.import_module(f"ddnerf_b200.models.{name}")
        setattr(pkg, name, sys.modules[f"models.{name}"])
    ours = importlib.import_module("ddnerf_b200.general_utils")
    try:
        theirs = importlib.import_module("general_utils")
        if theirs is ours or getattr(theirs, "__name__", "") == "ddnerf_b200.general_utils":
            raise ImportError
    except Exception:
        theirs = None
    if theirs is None:
        sys.modules["general_utils"] = ours
        sys.modules["general_utils.cfgnode"] = importlib.import_module("ddnerf_b200.general_utils.cfgnode")
        target = ours
    else:
        target = theirs
    for name in _UTIL_MODULES:
        mod = importlib.import_module(f"ddnerf_b200.general_utils.{name}")
        sys.modules[f"general_utils.{name}"] = mod
        setattr(target, name, mod)
        if target is not ours:                       # the reference's __init__ star-re-exports these modules
            for k in getattr(mod, "__all__", [k for k in vars(mod) if not k.startswith("_")]):
                setattr(target, k, getattr(mod, k))
    return sys.modules["models"]
