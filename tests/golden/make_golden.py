"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference (dadonda89/DDNeRF) ships no tests or fixtures, so the
vectors are outputs of the reference's own functions on seeded inputs, with every internal
``torch.rand`` / ``torch.randn`` draw recorded.  They pin ``oracle/ddnerf_oracle.py`` (CPU suite)
and the CUDA path (GPU suite).  Files are small ``.npz`` archives; nothing from the reference's
source is copied.
"""
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")

from general_utils.cfgnode import CfgNode  # noqa: E402  (reference)
from general_utils import math_utils as ref_math  # noqa: E402
from general_utils import nerf_helpers as ref_helpers  # noqa: E402
from general_utils.volume_rendering_utils import volume_render_radiance_field as ref_render  # noqa: E402
from models import samplers as ref_samplers  # noqa: E402
from models import dd_utils as ref_dd  # noqa: E402
from models import models as ref_models  # noqa: E402
from models import base_architectures as ref_arch  # noqa: E402

from oracle import ddnerf_oracle as orc  # noqa: E402  (only for init_mlp_params / ray synthesis)
from tests.synth import synth_rays, peaked_weights  # noqa: E402


class Recorder:
    """Records every torch.rand / torch.randn draw made while active."""

    def __init__(self):
        self.draws = []

    def __enter__(self):
        self._rand, self._randn = torch.rand, torch.randn

        def rand(*a, **k):
            out = self._rand(*a, **k)
            self.draws.append(("rand", out.clone()))
            return out

        def randn(*a, **k):
            out = self._randn(*a, **k)
            self.draws.append(("randn", out.clone()))
            return out

        torch.rand, torch.randn = rand, randn
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randn = self._rand, self._randn


def load_cfg(name):
    with open(f"/root/reference/configs/{name}") as f:
        return CfgNode(yaml.load(f, Loader=yaml.FullLoader))


def np_(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np_(v) for k, v in arrs.items() if v is not None})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def stage_samplers():
    torch.manual_seed(100)
    cfg = load_cfg("config_blender.yml")
    N = 24
    near = torch.full((N, 1), 2.0)
    far = torch.full((N, 1), 6.0)
    out = {}
    for tag, perturb, lindisp, nc in (("det", False, False, 16), ("jit", True, False, 32), ("lind", True, True, 8)):
        cfg.nerf.train.perturb = perturb
        cfg.nerf.train.lindisp = lindisp
        cfg.nerf.train.num_coarse = nc
        with Recorder() as r:
            t = ref_samplers.sample_first_cycle(cfg, near, far, "train")
        out[f"fc_{tag}_t"] = t
        if perturb:
            out[f"fc_{tag}_rand"] = r.draws[0][1]
    save("first_cycle", near=near, far=far, **out)

    for S, n in ((16, 17), (32, 33), (48, 25)):
        cfg.nerf.train.perturb = True
        cfg.nerf.train.lindisp = False
        cfg.nerf.train.num_coarse = S
        bins = ref_samplers.sample_first_cycle(cfg, near, far, "train")
        w_uniform = torch.rand(N, S) * 0.05
        w_peaked = peaked_weights(N, S, seed=7 + S)
        mus = torch.rand(N, S)
        sigmas = torch.rand(N, S) * 0.5 + 0.001
        lt = ref_math.approximate_cdf((0 - mus) / sigmas)
        pin = ref_math.approximate_cdf((1 - mus) / sigmas) - lt
        arrs = dict(bins=bins, w_uniform=w_uniform, w_peaked=w_peaked, mus=mus, sigmas=sigmas, lt=lt, pin=pin)
        for wname, w in (("uniform", w_uniform), ("peaked", w_peaked)):
            for pad in (True, False):
                cfg.train_params.pdf_padding = pad
                for det in (True, False):
                    key = f"{wname}_pad{int(pad)}_det{int(det)}"
                    with Recorder() as r:
                        s = ref_samplers.sample_pdf(bins, w, n, cfg, det=det)
                    arrs["mip_" + key] = s
                    if not det:
                        arrs["mip_" + key + "_rand"] = r.draws[0][1]
                    with Recorder() as r:
                        s = ref_samplers.sample_pdf_with_mu_sigma(bins, w, mus, sigmas, pin, lt, n, cfg, det=det)
                    arrs["dd_" + key] = s
                    if not det:
                        arrs["dd_" + key + "_rand"] = r.draws[0][1]
        save(f"resample_S{S}", near_cfg=cfg.dataset.near, far_cfg=cfg.dataset.far, **arrs)

    # one-cell special case, samplers.py:185-190
    bins = torch.cat((near, far), -1)
    w = torch.rand(N, 1)
    mus, sigmas = torch.rand(N, 1), torch.rand(N, 1) * 0.5 + 0.001
    lt = ref_math.approximate_cdf((0 - mus) / sigmas)
    pin = ref_math.approximate_cdf((1 - mus) / sigmas) - lt
    cfg.train_params.pdf_padding = True
    s = ref_samplers.sample_pdf_with_mu_sigma(bins, w, mus, sigmas, pin, lt, 9, cfg, det=True)
    save("resample_onecell", bins=bins, w=w, mus=mus, sigmas=sigmas, lt=lt, pin=pin, dd=s,
         near_cfg=cfg.dataset.near, far_cfg=cfg.dataset.far)


def stage_encoding():
    torch.manual_seed(101)
    for kind in ("blender", "ff", "360"):
        ro, rd, rad, near, far = synth_rays(kind, 16, seed=3)
        S = 16
        cfgname = {"blender": "config_blender.yml", "ff": "config_ff.yml", "360": "config_360.yml"}[kind]
        cfg = load_cfg(cfgname)
        cfg.nerf.train.num_coarse = S
        cfg.nerf.train.perturb = True
        nr = torch.full((ro.shape[0], 1), float(near))
        fr = torch.full((ro.shape[0], 1), float(far))
        t = ref_samplers.sample_first_cycle(cfg, nr, fr, "train")
        arrs = dict(ro=ro, rd=rd, rad=rad, t=t)
        for shape in ("cone", "cylinder"):
            means, covs = ref_math.cast_rays(t, ro, rd, rad, shape)
            arrs[f"means_{shape}"] = means
            arrs[f"covs_{shape}"] = covs
            arrs[f"ipe_{shape}"] = ref_math.integrated_pos_enc((means, covs))
        vd = rd / rd.norm(p=2, dim=-1, keepdim=True)
        arrs["dir_enc"] = ref_helpers.positional_encoding(vd, 4, True, True)
        save(f"encoding_{kind}", **arrs)


def stage_mlp():
    torch.manual_seed(102)
    x = torch.randn(96, 123) * 0.7
    for depth in (False, True):
        net = (ref_arch.DepthMipNeRFModel if depth else ref_arch.MipNeRFModel)(
            hidden_size=256, max_ipe_deg=16, num_encoding_fn_dir=4, include_input_xyz=False,
            include_input_dir=True, use_viewdirs=True)
        net.load_state_dict(orc.init_mlp_params(depth, seed=11 + int(depth)))
        xg = x.clone()
        y = net(xg)
        gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(5))
        y.backward(gy)
        grads = {"g_" + k: p.grad for k, p in net.named_parameters()
                 if p.numel() <= 256 * 96}
        save("mlp_depth" if depth else "mlp_plain", x=x, y=y, gy=gy, **grads)


def stage_render():
    torch.manual_seed(103)
    cfg_b = load_cfg("config_blender.yml")
    cfg_r = load_cfg("config_360.yml")
    N, S = 20, 32
    ro, rd, rad, near, far = synth_rays("blender", N, seed=9)
    cfg_b.nerf.train.num_coarse = S
    t = ref_samplers.sample_first_cycle(cfg_b, torch.full((N, 1), 2.0), torch.full((N, 1), 6.0), "train")
    raw = torch.randn(N, S, 4)
    raw[..., 3] = raw[..., 3] * 6 + 2          # some opaque samples
    raw[:3, :, 3] = -200.0                     # rays that hit nothing (sum of weights ~ 0)
    mus = torch.rand(N, S)
    arrs = dict(t=t, rd=rd, raw=raw, mus=mus)
    for tag, cfg, std, white, use_mus in (("blender", cfg_b, 0.0, False, False),
                                          ("blender_noise_white", cfg_b, 1.0, True, False),
                                          ("blender_mus", cfg_b, 1.0, False, True),
                                          ("real_mus", cfg_r, 0.5, False, True),
                                          ("nocfg", None, 0.0, False, False)):
        rawg = raw.clone().requires_grad_(True)
        musg = mus.clone().requires_grad_(True)
        with Recorder() as r:
            outs = ref_render(rawg, t, rd, radiance_field_noise_std=std, white_background=white,
                              mus=musg if use_mus else None, cfg=cfg)
        if std > 0:
            arrs[f"{tag}_randn"] = r.draws[0][1]
        names = ("rgb_map", "disp", "acc", "weights", "depth", "cdisp", "rgb")
        for nme, o in zip(names, outs):
            arrs[f"{tag}_{nme}"] = o
        # a loss touching every differentiable output, fixed cotangents
        g = torch.Generator().manual_seed(17)
        loss = 0
        for nme, o in zip(names[:6], outs[:6]):
            if o is None:
                continue
            ct = torch.randn(o.shape, generator=g)
            if nme in ("disp", "cdisp"):
                ct = ct * 1e-2
            arrs[f"{tag}_ct_{nme}"] = ct
            loss = loss + (o * ct).sum()
        loss.backward()
        arrs[f"{tag}_g_raw"] = rawg.grad
        if use_mus:
            arrs[f"{tag}_g_mus"] = musg.grad
    save("render", **arrs)


def stage_dp_loss():
    torch.manual_seed(104)
    cfg_b = load_cfg("config_blender.yml")
    cfg_r = load_cfg("config_360.yml")
    N = 28
    for S0, S1 in ((16, 16), (32, 32), (24, 40)):
        cfg_b.nerf.train.num_coarse = S0
        near = torch.full((N, 1), 2.0)
        far = torch.full((N, 1), 6.0)
        t0 = ref_samplers.sample_first_cycle(cfg_b, near, far, "train")
        arrs = dict(t0=t0)
        for wname in ("uniform", "peaked", "bumpy"):
            # "peaked" has runs of exact zeros (cdf ties; KL gradients there are ill-conditioned, ~p1/q
            # with q ~ 1e-12, so only the loss is pinned); "bumpy" is peaked with a floor (gradients pinned)
            w0 = (torch.rand(N, S0) * 0.05 if wname == "uniform" else
                  peaked_weights(N, S0, seed=S0) + (0.002 if wname == "bumpy" else 0.0))
            mus = torch.rand(N, S0)
            sigmas = torch.rand(N, S0) * 0.5 + 0.001
            lt = ref_math.approximate_cdf((0 - mus) / sigmas)
            pin = ref_math.approximate_cdf((1 - mus) / sigmas) - lt
            cfg_b.train_params.pdf_padding = True
            t1 = ref_samplers.sample_pdf_with_mu_sigma(t0, w0, mus, sigmas * 1.7,
                                                       ref_math.approximate_cdf((1 - mus) / (sigmas * 1.7)) -
                                                       ref_math.approximate_cdf((0 - mus) / (sigmas * 1.7)),
                                                       ref_math.approximate_cdf((0 - mus) / (sigmas * 1.7)),
                                                       S1 + 1, cfg_b, det=False).detach()
            w1 = (torch.rand(N, S1) * 0.05 if wname == "uniform" else
                  peaked_weights(N, S1, seed=S1 + 1) + (0.002 if wname == "bumpy" else 0.0))
            w1[:, -1] += 1e-10
            arrs.update({f"{wname}_w0": w0, f"{wname}_mus": mus, f"{wname}_sigmas": sigmas,
                         f"{wname}_lt": lt, f"{wname}_pin": pin, f"{wname}_t1": t1, f"{wname}_w1": w1})
            for cname, cfg in (("blender", cfg_b), ("real", cfg_r)):
                w0g = w0.clone().requires_grad_(True)
                mg = mus.clone().requires_grad_(True)
                sg = sigmas.clone().requires_grad_(True)
                loss = ref_dd.estimate_dp_loss(t1, t0, w1, w0g, mg, sg, lt, pin, cfg)
                loss.backward()
                arrs.update({f"{wname}_{cname}_loss": loss, f"{wname}_{cname}_g_w0": w0g.grad,
                             f"{wname}_{cname}_g_mus": mg.grad, f"{wname}_{cname}_g_sigmas": sg.grad})
        save(f"dp_loss_{S0}_{S1}", **arrs)


def stage_rays():
    """f1: get_ray_bundle (nerf_helpers.py:67-125) and ndc_mipnerf_rays (dataset_helpers.py:3-42) of the reference on
    small frames: a rotated + translated pinhole pose, and a forward-facing pose through the NDC projection (even
    width, identity rotation -> an exactly-zero direction component, the epsilon branch of :114-115)."""
    from data_utils.dataset_helpers import ndc_mipnerf_rays as ref_ndc
    from ddnerf_b200.rays import pose_spherical
    out = {}
    H, W, focal = 13, 18, 21.37
    pose = pose_spherical(37.0, -25.0, 3.3)
    ro, rd, rad = ref_helpers.get_ray_bundle(H, W, focal, pose)
    out.update(p_H=H, p_W=W, p_focal=focal, p_pose=pose, p_ro=ro.clone(), p_rd=rd, p_rad=rad)
    H, W, focal = 14, 20, 17.5
    pose = torch.eye(4)
    pose[:3, 3] = torch.tensor([0.11, -0.07, 0.02])
    ro, rd, _ = ref_helpers.get_ray_bundle(H, W, focal, pose)
    o, d, r = ref_ndc(H, W, focal, ro.clone(), rd, near=1)
    out.update(n_H=H, n_W=W, n_focal=focal, n_pose=pose, n_ro=o, n_rd=d, n_rad=r, n_rd_cam=rd)
    save("ray_bundle", **{k: (torch.tensor(v) if not isinstance(v, torch.Tensor) else v) for k, v in out.items()})


def stage_frame():
    """f4: cast_to_image / cast_to_disparity_image of the reference (validation_utils/visualization.py:11-27) and the video
    frame assembly of its render loop (render_video.py:96-101, executed from the reference's own source text) on a small
    synthetic frame whose colours cover the renderer's full range [-0.001, 1.001] (volume_rendering_utils.py:25-27).
    ``matplotlib`` and ``imageio`` are not installed here; the module only uses them for plots / file output, so empty
    stand-ins are registered before the import."""
    import textwrap
    import types
    for name in ("matplotlib", "matplotlib.pyplot", "imageio"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    from validation_utils import visualization as ref_vis
    import cv2
    g = torch.Generator().manual_seed(11)
    H, W = 23, 37
    rgb = torch.rand(H, W, 3, generator=g) * 1.002 - 0.001
    rgb[0, 0] = torch.tensor([-0.001, 1.001, 0.5])
    rgb[0, 1] = torch.tensor([1.0, 0.0, 254.5 / 255.0])
    rgb[0, 2] = torch.tensor([1.0 / 255.0, 2.0 / 255.0 - 1e-7, 0.99999])
    disp = 1.0 / (torch.rand(H, W, generator=g) * 4.0 + 0.2)
    img = ref_vis.cast_to_image(rgb[..., :3])                       # [3,H,W] uint8
    d8 = ref_vis.cast_to_disparity_image(disp)                       # [1,H,W] uint8
    # render_video.py:78-101: the frame written to the video (disp = cast_to_disparity_image(disp).squeeze())
    lines = open("/root/reference/render_video.py").read().splitlines()[97:103]
    assert lines[0].strip().startswith("rgb = 255*rgb") and lines[-1].strip().startswith("frame = cv2.cvtColor"), lines
    scope = {"rgb": rgb.clone(), "disp": d8.squeeze(), "torch": torch, "np": np, "cv2": cv2}
    exec(textwrap.dedent("\n".join(lines)), scope)
    save("frame_post", rgb=rgb, disp=disp, rgb8_chw=img, disp8=d8, video_bgr=scope["frame"])


def subsample_grads(prefix, module):
    out = {}
    for k, p in module.named_parameters():
        g = p.grad
        out[f"{prefix}{k}"] = g if g.numel() <= 4096 else g.flatten()[::97]
    return out


def end_to_end():
    runs = (("dd_blender_train", "config_blender.yml", "train", 16, 16, 40),
            ("dd_360_train", "config_360.yml", "train", 32, 32, 24),
            ("dd_ff_val", "config_ff.yml", "validation", 16, 16, 48),
            ("mip_blender_train", "config_blender_mipnerf.yml", "train", 32, 32, 24),
            ("mip_blender_val", "config_blender_mipnerf.yml", "validation", 16, 24, 32))
    for tag, cfgname, mode, nc, nf, N in runs:
        torch.manual_seed(200)
        cfg = load_cfg(cfgname)
        kind = "ff" if "ff" in cfgname else ("360" if "360" in cfgname else "blender")
        ro, rd, rad, near, far = synth_rays(kind, N, seed=21)
        cfg.dataset.near, cfg.dataset.far = float(near), float(far)
        cfg.nerf[mode].num_coarse, cfg.nerf[mode].num_fine = nc, nf
        cfg.train_params.dist_reg_coeficient = min(max(1 / nc, 0.01), 0.12)     # train_model.py:124-125
        if mode == "validation" and cfg.nerf.type == "DDNerfModel":            # render_video.py:40-42
            cfg.train_params.pdf_padding = False
            cfg.train_params.gaussian_smooth_factor = cfg.train_params.final_smooth
        model = getattr(ref_models, cfg.nerf.type)(cfg)
        is_dd = cfg.nerf.type == "DDNerfModel"
        model.coarse.load_state_dict(orc.init_mlp_params(is_dd, seed=31))
        if is_dd:
            model.fine.load_state_dict(orc.init_mlp_params(False, seed=32))
        target = torch.rand(N, 3, generator=torch.Generator().manual_seed(77))
        arrs = dict(ro=ro, rd=rd, rad=rad, near=near, far=far, target=target)
        with Recorder() as r:
            if mode == "train":
                out = model.run_iter(ro, rd, rad, mode="train", rgb_target=target)
                loss = 0
                for j in range(2):
                    loss = loss + cfg.train_params.loss_coeficients[j] * torch.nn.functional.mse_loss(out[j]["rgb"], target)
                if is_dd:
                    loss = loss + cfg.train_params.dp_coeficient * out[1]["dp_loss"].mean()
                loss.backward()
                arrs["loss"] = loss
                arrs.update(subsample_grads("gc_", model.coarse))
                if is_dd:
                    arrs.update(subsample_grads("gf_", model.fine))
            else:
                with torch.no_grad():
                    out = model.run_iter(ro.view(N // 8, 8, 3), rd.view(N // 8, 8, 3), rad.view(N // 8, 8, 1),
                                         mode="validation")
        for i, (kindr, d) in enumerate(r.draws):
            arrs[f"draw{i}_{kindr}"] = d
        for j in range(2):
            for k, v in out[j].items():
                if isinstance(v, torch.Tensor):
                    arrs[f"out{j}_{k}"] = v
        save("e2e_" + tag, **arrs)


if __name__ == "__main__":
    if "--rays-only" in sys.argv:
        stage_rays()
        sys.exit(0)
    if "--frame-only" in sys.argv:
        stage_frame()
        sys.exit(0)
    stage_samplers()
    stage_encoding()
    stage_mlp()
    stage_render()
    stage_dp_loss()
    stage_rays()
    stage_frame()
    end_to_end()
