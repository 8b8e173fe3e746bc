"""TEST / BASELINE INFRASTRUCTURE -- locates and imports the UNMODIFIED reference (dadonda89/DDNeRF).

The reference is pure Python.  ``__graft_entry__.build()`` copies its tree, as is, from ``/root/reference`` to
``baseline/_ref/`` in the authoring container (git-ignored: it never enters this repository's history, but it travels
to the GPU box with the working tree, where ``/root/reference`` does not exist).  Only ``tests/``, ``bench.py``'s
reference / cpu_baseline legs and ``tests/golden/make_golden.py`` import this module; nothing under ``ddnerf_b200/``
does.
"""
import importlib
import os
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(REPO, "baseline", "_ref"), "/root/reference")


def reference_root():
    """Directory of the reference tree, or None."""
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "models", "models.py")) and os.path.isfile(os.path.join(c, "train_model.py")):
            return c
    return None


def stub_missing_driver_deps():
    """``imageio``, ``matplotlib``, ``scikit-image`` and ``lpips`` are not installed in this image.  The reference's data / plot
    modules import them at module level; the hot path never calls them.  Register minimal stand-ins (``imageio.imread``
    through PIL so the Blender loader works) unless the real packages exist."""
    def have(name):
        try:
            importlib.import_module(name)
            return True
        except Exception:
            return False
    if not have("imageio"):
        import numpy as np
        from PIL import Image
        m = types.ModuleType("imageio")
        m.imread = lambda f, *a, **k: np.array(Image.open(f))
        m.imwrite = lambda f, arr, *a, **k: Image.fromarray(np.asarray(arr)).save(f)
        sys.modules["imageio"] = m
    if not have("matplotlib"):
        m, p = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        m.pyplot = p
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = m, p
    if not have("skimage"):
        m = types.ModuleType("skimage")
        for sub in ("transform", "measure", "metrics"):
            s = types.ModuleType(f"skimage.{sub}")
            setattr(m, sub, s)
            sys.modules[f"skimage.{sub}"] = s
        # (metrics of eval_nerf.py, outside the hot path: placeholders that keep the driver running)
        sys.modules["skimage.measure"].compare_ssim = lambda a, b, full=False, **k: (0.0, None) if full else 0.0
        sys.modules["skimage.metrics"].structural_similarity = lambda a, b, **k: 0.0
        sys.modules["skimage"] = m
    if not have("lpips"):
        import torch
        m = types.ModuleType("lpips")

        class LPIPS:                      # eval_nerf.py:87: the perceptual metric needs downloaded AlexNet weights
            def __init__(self, *a, **k):
                pass

            def __call__(self, x, y):
                return torch.zeros(1, 1, 1, 1)
        m.LPIPS = LPIPS
        sys.modules["lpips"] = m


def import_reference():
    """Put the reference tree first on ``sys.path`` and import its hot-path packages under their own top-level names
    (``models``, ``general_utils``).  Returns (root, models.models module, CfgNode).  Must not be combined with
    ``ddnerf_b200.install_as_reference()`` in one process."""
    root = reference_root()
    if root is None:
        raise RuntimeError("the reference tree is not available (neither baseline/_ref nor /root/reference)")
    mod = sys.modules.get("models")
    if mod is not None and getattr(mod, "__name__", "").startswith("ddnerf_b200"):
        raise RuntimeError("ddnerf_b200.install_as_reference() is active in this process")
    if root not in sys.path:
        sys.path.insert(0, root)
    ref_models = importlib.import_module("models.models")
    cfgnode = importlib.import_module("general_utils.cfgnode")
    return root, ref_models, cfgnode.CfgNode


def load_reference_cfg(name):
    """One of the reference's shipped YAML files (configs/<name>.yml) as its own CfgNode."""
    import yaml
    root, _, CfgNode = import_reference()
    with open(os.path.join(root, "configs", name + ".yml")) as f:
        return CfgNode(yaml.load(f, Loader=yaml.FullLoader))
