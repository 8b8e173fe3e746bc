"""The drop-in contract of SURVEY.md section 8b / BASELINE.json north_star: the reference's OWN drivers run unchanged on
top of this package.  ``train_model.py`` of the unmodified reference tree (baseline/_ref, copied there by
``__graft_entry__.build()``; /root/reference in the authoring container) is executed for three iterations -- dataset
loading, model construction through ``getattr(models, cfg.nerf.type)(cfg)``, its two ``torch.optim.Adam`` objects, the
per-iteration mutation of ``cfg.train_params``, validation renders into TensorBoard, the checkpoint -- with
``ddnerf_b200.install_as_reference()`` as the only addition, on a tiny synthetic Blender-format scene."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make_blender_scene(root, n_train=3, size=24):
    """transforms_{train,val,test}.json + RGBA PNGs in the layout data_utils/load_blender.py:66-112 reads."""
    from PIL import Image
    from ddnerf_b200.rays import pose_spherical
    rng = np.random.default_rng(0)
    for split, n in (("train", n_train), ("val", 2), ("test", 1)):
        os.makedirs(os.path.join(root, split), exist_ok=True)
        frames = []
        for i in range(n):
            img = rng.integers(0, 256, size=(size, size, 4), dtype=np.uint8)
            img[..., 3] = 255
            img[: size // 3, :, 3] = 0                               # transparent band: alpha-composited background
            Image.fromarray(img, "RGBA").save(os.path.join(root, split, f"r_{i}.png"))
            pose = pose_spherical(40.0 * i + 10.0 * len(split), -30.0, 4.0).numpy().tolist()
            frames.append({"file_path": f"./{split}/r_{i}", "transform_matrix": pose})
        with open(os.path.join(root, f"transforms_{split}.json"), "w") as f:
            json.dump({"camera_angle_x": 0.6911112070083618, "frames": frames}, f)


_LAUNCHER = r"""
import os, sys, runpy
repo, ref, script = sys.argv[1], sys.argv[2], sys.argv[3]
script_args = sys.argv[4:]
sys.path.insert(0, repo)
from oracle import reference_loader as RL            # test infrastructure: stubs for imageio / matplotlib / skimage
RL.stub_missing_driver_deps()
sys.path.insert(0, ref)                                # the reference tree: train_model.py, data_utils, validation_utils
import ddnerf_b200
ddnerf_b200.install_as_reference()                     # <- the one line a maintainer adds (INTEGRATION.md section 2)
from models import models
assert models.DDNerfModel.__module__ == "ddnerf_b200.models.models"
from ddnerf_b200 import _lib
before = _lib.load().ddnerf_launch_count()
sys.argv = [os.path.join(ref, script)] + script_args
runpy.run_path(sys.argv[0], run_name="__main__")
print("DDNERF_KERNEL_LAUNCHES", _lib.load().ddnerf_launch_count() - before)
"""


@pytest.mark.parametrize("cfg_name,mlp_mode", [("config_blender", "fp32"), ("config_blender_mipnerf", "bf16")])
def test_reference_train_model_runs_unchanged(tmp_path, cfg_name, mlp_mode):
    import yaml
    sys.path.insert(0, REPO)
    from oracle import reference_loader as RL
    ref = RL.reference_root()
    if ref is None:
        pytest.skip("reference tree not available (baseline/_ref is created by __graft_entry__.build() where /root/reference exists)")
    scene = tmp_path / "scene"
    _make_blender_scene(str(scene))
    with open(os.path.join(ref, "configs", cfg_name + ".yml")) as f:
        cfg = yaml.load(f, Loader=yaml.FullLoader)
    cfg["experiment"].update(id="drivers_unchanged", logdir=str(tmp_path / "logs"), train_iters=3, validate_every=2,
                             save_every=2, print_every=1)
    cfg["dataset"].update(basedir=str(scene), half_res=False, testskip=1)
    cfg["nerf"]["train"].update(num_random_rays=192, num_coarse=16, num_fine=16)
    cfg["nerf"]["validation"].update(num_coarse=16, num_fine=16)
    cfg_path = tmp_path / "config.yml"
    with open(cfg_path, "w") as f:
        yaml.safe_dump(cfg, f)
    launcher = tmp_path / "launch.py"
    launcher.write_text(_LAUNCHER)
    env = dict(os.environ, DDNERF_MLP_MODE=mlp_mode)
    r = subprocess.run([sys.executable, str(launcher), REPO, ref, "train_model.py", "--config", str(cfg_path)], capture_output=True,
                       text=True, cwd=str(tmp_path), env=env, timeout=900)
    tail = (r.stdout[-3000:] + "\n--- stderr ---\n" + r.stderr[-3000:])
    assert r.returncode == 0, tail
    assert "Done!" in r.stdout, tail                                                    # train_model.py:264
    launches = [int(l.split()[1]) for l in r.stdout.splitlines() if l.startswith("DDNERF_KERNEL_LAUNCHES")]
    assert launches and launches[0] > 50, tail                                          # the CUDA path did the work
    logdir = tmp_path / "logs" / "drivers_unchanged"
    ck = torch.load(logdir / "checkpoint.ckpt", map_location="cpu", weights_only=False)  # train_model.py:248-263
    assert ck["iter"] == 2 and "optimizer_1_state_dict" in ck
    names = list(ck["model_1_state_dict"])
    assert names[0] == "layers_xyz.0.weight" and ck["model_1_state_dict"]["layers_xyz.5.weight"].shape == (256, 352)
    if cfg_name == "config_blender":
        assert "model_2_state_dict" in ck and "fc_mu_sigma.weight" in names
    for v in ck["model_1_state_dict"].values():
        assert torch.isfinite(v).all()
    assert any(f.startswith("events.out.tfevents") for f in os.listdir(logdir))          # the Documenter wrote its scalars
    assert os.path.exists(logdir / "config.yml")

    # resume from that checkpoint with the reference's own flag (train_model.py:77-81,110-118)
    cfg["experiment"]["train_iters"] = 5
    with open(cfg_path, "w") as f:
        yaml.safe_dump(cfg, f)
    r2 = subprocess.run([sys.executable, str(launcher), REPO, ref, "train_model.py", "--config", str(cfg_path), "--load-checkpoint",
                         str(logdir / "checkpoint.ckpt")], capture_output=True, text=True, cwd=str(tmp_path), env=env, timeout=900)
    assert r2.returncode == 0 and "Done!" in r2.stdout, r2.stdout[-2000:] + r2.stderr[-3000:]
    ck2 = torch.load(logdir / "checkpoint.ckpt", map_location="cpu", weights_only=False)
    assert ck2["iter"] == 4


    # the reference's render driver on that checkpoint (render_video.py:17-112): config.yml + checkpoint.ckpt from the log
    # directory, load_weights_from_checkpoint, one validation-mode run_iter per render pose, cast_to_disparity_image,
    # frames into its cv2 video writer and (--save_images) PNGs through imageio
    r3 = subprocess.run([sys.executable, str(launcher), REPO, ref, "render_video.py", "--logdir", str(logdir), "--save_images", "1"],
                        capture_output=True, text=True, cwd=str(tmp_path), env=env, timeout=900)
    tail3 = r3.stdout[-2000:] + "\n--- stderr ---\n" + r3.stderr[-3000:]
    assert r3.returncode == 0, tail3
    launches = [int(l.split()[1]) for l in r3.stdout.splitlines() if l.startswith("DDNERF_KERNEL_LAUNCHES")]
    assert launches and launches[0] > 50, tail3
    pngs = sorted(os.listdir(logdir / "video" / "images"))
    assert len(pngs) >= 2 and len(os.listdir(logdir / "video" / "disparity")) == len(pngs), tail3
    from PIL import Image
    img = np.array(Image.open(logdir / "video" / "images" / pngs[0]))
    assert img.shape[:2] == (24, 24) and img.dtype == np.uint8

    # ... and its evaluation driver (eval_nerf.py:20-166): validation-mode run_iter with rgb_target on every validation
    # pose, save_validation_images over both output dicts, PSNR from the returned rgb (the perceptual / SSIM metrics are
    # placeholders here: lpips and scikit-image are not installed)
    r4 = subprocess.run([sys.executable, str(launcher), REPO, ref, "eval_nerf.py", "--logdir", str(logdir)],
                        capture_output=True, text=True, cwd=str(tmp_path), env=env, timeout=900)
    tail4 = r4.stdout[-2000:] + "\n--- stderr ---\n" + r4.stderr[-3000:]
    assert r4.returncode == 0, tail4
    val = logdir / "validation"
    assert os.path.exists(val / "results.txt") and os.path.exists(val / "val_image_1" / "rgb_fine.png"), tail4
    assert "psnr_fine" in open(val / "results.txt").read()
