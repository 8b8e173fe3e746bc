"""CPU suite: pins oracle/ddnerf_oracle.py against vectors produced by the real reference
(tests/golden/make_golden.py).  Index outputs must be bit-exact; floats within 1e-6 abs/rel
(the oracle reorders nothing, differences are last-ulp effects of fused ops)."""
import numpy as np
import pytest
import torch

from oracle import ddnerf_oracle as orc
from tests.conftest import load_golden

TOL = dict(rtol=2e-6, atol=2e-6)


def close(a, b, **kw):
    kw = {**TOL, **kw}
    torch.testing.assert_close(a, b, equal_nan=True, **kw)


def test_first_cycle():
    g = load_golden("first_cycle")
    close(orc.sample_first_cycle(g["near"], g["far"], 16), g["fc_det_t"].expand(24, 17))
    close(orc.sample_first_cycle(g["near"], g["far"], 32, False, g["fc_jit_rand"]), g["fc_jit_t"])
    close(orc.sample_first_cycle(g["near"], g["far"], 8, True, g["fc_lind_rand"]), g["fc_lind_t"])


@pytest.mark.parametrize("S", [16, 32, 48])
@pytest.mark.parametrize("wname", ["uniform", "peaked"])
@pytest.mark.parametrize("pad", [True, False])
@pytest.mark.parametrize("det", [True, False])
def test_resamplers(S, wname, pad, det):
    g = load_golden(f"resample_S{S}")
    key = f"{wname}_pad{int(pad)}_det{int(det)}"
    w = g["w_" + wname]
    n = g["mip_" + key].shape[1]
    rand = None if det else g["mip_" + key + "_rand"]
    s, _ = orc.sample_pdf(g["bins"], w, n, pad, rand)
    close(s, g["mip_" + key])
    rand = None if det else g["dd_" + key + "_rand"]
    s, _ = orc.sample_pdf_with_mu_sigma(g["bins"], w, g["mus"], g["sigmas"], g["pin"], g["lt"], n, pad,
                                        float(g["near_cfg"]), float(g["far_cfg"]), rand)
    close(s, g["dd_" + key], rtol=1e-5, atol=1e-5)


def test_resampler_one_cell():
    g = load_golden("resample_onecell")
    s, _ = orc.sample_pdf_with_mu_sigma(g["bins"], g["w"], g["mus"], g["sigmas"], g["pin"], g["lt"], 9, True,
                                        float(g["near_cfg"]), float(g["far_cfg"]))
    close(s, g["dd"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("kind", ["blender", "ff", "360"])
def test_encoding(kind):
    g = load_golden(f"encoding_{kind}")
    for shape in ("cone", "cylinder"):
        means, covs = orc.cast_rays(g["t"], g["ro"], g["rd"], g["rad"], shape)
        close(means, g[f"means_{shape}"])
        close(covs, g[f"covs_{shape}"])
        close(orc.integrated_pos_enc(means, covs), g[f"ipe_{shape}"])
    vd = g["rd"] / g["rd"].norm(p=2, dim=-1, keepdim=True)
    close(orc.positional_encoding(vd), g["dir_enc"])


@pytest.mark.parametrize("depth", [False, True])
def test_mlp(depth):
    g = load_golden("mlp_depth" if depth else "mlp_plain")
    params = {k: v.requires_grad_(True) for k, v in orc.init_mlp_params(depth, seed=11 + int(depth)).items()}
    y = orc.mlp_forward(params, g["x"])
    close(y, g["y"], rtol=1e-5, atol=1e-5)
    y.backward(g["gy"])
    for k, p in params.items():
        if "g_" + k in g:
            close(p.grad, g["g_" + k], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("tag,std,white,use_mus,blender", [
    ("blender", 0.0, False, False, True), ("blender_noise_white", 1.0, True, False, True),
    ("blender_mus", 1.0, False, True, True), ("real_mus", 0.5, False, True, False),
    ("nocfg", 0.0, False, False, False)])
def test_render(tag, std, white, use_mus, blender):
    g = load_golden("render")
    raw = g["raw"].clone().requires_grad_(True)
    mus = g["mus"].clone().requires_grad_(True)
    noise = g[f"{tag}_randn"] * std if std > 0 else None
    outs = orc.volume_render(raw, g["t"], g["rd"], noise, white, blender, mus if use_mus else None)
    names = ("rgb_map", "disp", "acc", "weights", "depth", "cdisp", "rgb")
    loss = 0
    for nme, o in zip(names, outs):
        if o is None:
            assert f"{tag}_{nme}" not in g
            continue
        close(o, g[f"{tag}_{nme}"], rtol=1e-5, atol=1e-6)
        if f"{tag}_ct_{nme}" in g:
            loss = loss + (o * g[f"{tag}_ct_{nme}"]).sum()
    loss.backward()
    close(raw.grad, g[f"{tag}_g_raw"], rtol=1e-4, atol=1e-5)
    if use_mus:
        close(mus.grad, g[f"{tag}_g_mus"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("shape", ["16_16", "32_32", "24_40"])
@pytest.mark.parametrize("wname", ["uniform", "peaked", "bumpy"])
@pytest.mark.parametrize("cname", ["blender", "real"])
def test_dp_loss(shape, wname, cname):
    g = load_golden(f"dp_loss_{shape}")
    w0 = g[f"{wname}_w0"].clone().requires_grad_(True)
    mus = g[f"{wname}_mus"].clone().requires_grad_(True)
    sig = g[f"{wname}_sigmas"].clone().requires_grad_(True)
    loss = orc.estimate_dp_loss(g[f"{wname}_t1"], g["t0"], g[f"{wname}_w1"], w0, mus, sig, g[f"{wname}_lt"],
                                g[f"{wname}_pin"], cname == "blender")
    close(loss, g[f"{wname}_{cname}_loss"], rtol=1e-5, atol=1e-7)
    if wname == "peaked":
        return      # gradients ~ p1/q with q ~ 1e-12 in empty space: 1-ulp noise dominates (see make_golden.py)
    loss.backward()
    close(w0.grad, g[f"{wname}_{cname}_g_w0"], rtol=1e-4, atol=1e-6)
    close(mus.grad, g[f"{wname}_{cname}_g_mus"], rtol=1e-4, atol=1e-6)
    close(sig.grad, g[f"{wname}_{cname}_g_sigmas"], rtol=1e-4, atol=1e-6)


E2E = {
    "dd_blender_train": dict(model="DDNerfModel", blender=True, nc=16, nf=16, train=True, smooth=1.7, pad=True),
    "dd_360_train": dict(model="DDNerfModel", blender=False, nc=32, nf=32, train=True, smooth=1.7, pad=True),
    "dd_ff_val": dict(model="DDNerfModel", blender=False, nc=16, nf=16, train=False, smooth=1.1, pad=False),
    "mip_blender_train": dict(model="GeneralMipNerfModel", blender=True, nc=32, nf=32, train=True, smooth=1.7,
                              pad=True, coefs=(1.0, 0.1)),
    "mip_blender_val": dict(model="GeneralMipNerfModel", blender=True, nc=16, nf=24, train=False, smooth=1.7,
                            pad=True, coefs=(1.0, 0.1)),
}


def e2e_setup(tag):
    """Shared by the CPU (oracle) and GPU (CUDA path) end-to-end tests."""
    spec = E2E[tag]
    g = load_golden("e2e_" + tag)
    cfg = orc.PathConfig(model=spec["model"], near=float(g["near"]), far=float(g["far"]), num_coarse=spec["nc"],
                         num_fine=spec["nf"], perturb=spec["train"], noise_std=1.0, blender=spec["blender"],
                         pdf_padding=spec["pad"], gaussian_smooth_factor=spec["smooth"],
                         dist_reg_coeficient=min(max(1 / spec["nc"], 0.01), 0.12),
                         loss_coeficients=spec.get("coefs", (1.0, 1.0)))
    is_dd = spec["model"] == "DDNerfModel"
    pc = orc.init_mlp_params(is_dd, seed=31)
    pf = orc.init_mlp_params(False, seed=32) if is_dd else None
    draws = [g[k] for k in sorted((k for k in g if k.startswith("draw")), key=lambda s: int(s[4:].split("_")[0]))]
    if spec["train"]:
        rnd = dict(t_rand=draws[0], noise0=draws[1], u_rand=draws[2], noise1=draws[3])
    else:
        rnd = dict(noise0=draws[0], noise1=draws[1])
    rays = orc.pack_rays(g["ro"], g["rd"], g["rad"], cfg.near, cfg.far)
    return spec, g, cfg, pc, pf, rnd, rays


def check_outputs(out, g, tol, flat=False):
    for j in range(2):
        for k in ("rgb", "disp", "acc", "weights", "depth", "corrected_disp_map", "dp_loss", "mus_loss", "sig_loss"):
            key = f"out{j}_{k}"
            if key not in g:
                continue
            ref = g[key]
            got = out[j][k]
            if k == "disp" or k == "corrected_disp_map":
                got, ref = 1.0 / got, 1.0 / ref      # compare in depth units (disp = 1/depth can be huge)
            close(got.reshape(ref.shape), ref, rtol=tol, atol=tol)


@pytest.mark.parametrize("tag", list(E2E))
def test_end_to_end(tag):
    spec, g, cfg, pc, pf, rnd, rays = e2e_setup(tag)
    if spec["train"]:
        loss, out, gc, gf = orc.train_step(cfg, pc, pf, rays, g["target"], rnd)
        close(loss, g["loss"], rtol=1e-5, atol=1e-6)
        for prefix, grads in (("gc_", gc), ("gf_", gf)):
            if grads is None:
                continue
            for k, v in grads.items():
                ref = g[prefix + k]
                got = v if v.numel() <= 4096 else v.flatten()[::97]
                close(got, ref, rtol=2e-3, atol=2e-6)
    else:
        with torch.no_grad():
            out = orc.predict_dd(cfg, pc, pf, rays, rnd) if pf is not None else orc.predict_mip(cfg, pc, rays, rnd)
    check_outputs(out, g, 2e-5)
    if pf is not None:
        for k in ("mus", "sigmas", "smoothed_sigmas"):
            close(out[0][k], g[f"out0_{k}"], rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------
# f1: the host-side restatement of the ray set-up (ddnerf_b200/rays.py) against the reference's own output
# ---------------------------------------------------------------------------------------------
def test_ray_bundle_restatement_golden():
    from ddnerf_b200 import rays as R
    g = load_golden("ray_bundle")
    H, W, focal = int(g["p_H"]), int(g["p_W"]), float(g["p_focal"])
    ro, rd, rad = R.get_ray_bundle(H, W, focal, g["p_pose"])
    for got, key in ((ro, "p_ro"), (rd, "p_rd"), (rad, "p_rad")):
        torch.testing.assert_close(got, g[key], rtol=1e-6, atol=1e-7)
    H, W, focal = int(g["n_H"]), int(g["n_W"]), float(g["n_focal"])
    ro, rd, _ = R.get_ray_bundle(H, W, focal, g["n_pose"])
    torch.testing.assert_close(rd, g["n_rd_cam"], rtol=1e-6, atol=1e-7)
    o, d, r = R.ndc_mipnerf_rays(H, W, focal, ro, rd, near=1)
    for got, key in ((o, "n_ro"), (d, "n_rd"), (r.squeeze(-1), "n_rad")):
        torch.testing.assert_close(got, g[key].reshape(got.shape), rtol=2e-6, atol=1e-6)


# ---------------------------------------------------------------------------------------------
# f4: frame post-processing against what the reference's own functions returned (tests/golden/frame_post.npz:
# validation_utils/visualization.py:11-27 and the frame assembly of render_video.py:98-103, generated with
# matplotlib / imageio stubbed -- make_golden.py: stage_frame)
# ---------------------------------------------------------------------------------------------
def test_frame_post_golden():
    g = load_golden("frame_post")
    rgb, disp = g["rgb"], g["disp"]
    rgb8 = orc.cast_to_image(rgb)                                   # [H,W,3]
    assert np.array_equal(np.moveaxis(rgb8, -1, 0), g["rgb8_chw"].numpy())
    d8 = orc.cast_to_disparity_image(disp)                           # [H,W]
    assert np.array_equal(d8[None], g["disp8"].numpy())
    assert np.array_equal(orc.video_frame(rgb, d8), g["video_bgr"].numpy())
