"""CPU oracle for the DDNeRF / mip-NeRF per-ray hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ddnerf_b200/`` imports this module.  The only
permitted users are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``; there it is the checker or the timed CPU baseline,
never the thing shipped.

It restates, in plain torch fp32 on the CPU, the algorithm of every function on the path that
starts at ``models/models.py:40`` (``run_iter``) of dadonda89/DDNeRF.  Every function cites the
reference file:line it follows.  Differences from the reference's *code* (not its results):

* every random draw (``torch.rand`` / ``torch.randn`` inside samplers.py:57,102,165 and
  volume_rendering_utils.py:31) is an explicit argument, so both sides of a parity test see
  the same numbers;
* the O(S^2) boolean-mask interval search (samplers.py:106-116, 175-193, dd_utils.py:43-52) is
  restated with ``torch.searchsorted`` plus the first-maximum tie rule of ``torch.max``; the
  golden vectors in ``tests/golden`` (made by ``tests/golden/make_golden.py`` from the real
  reference) pin that the two agree, indices included.

Parity status: PINNED against outputs of the reference itself (the reference ships no tests or
golden vectors of its own), see ``tests/test_oracle_golden.py``.
"""
import math
import warnings

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


# --------------------------------------------------------------------------------------------
# configuration record (what the hot path reads from cfg)
# --------------------------------------------------------------------------------------------
class PathConfig:
    """The cfg attributes the path reads (models.py:49,90-104,122,155-156; samplers.py:16-78;
    volume_rendering_utils.py:51; dd_utils.py:12), flattened into one record."""

    def __init__(self, model="DDNerfModel", near=2.0, far=6.0, num_coarse=32, num_fine=32,
                 perturb=True, lindisp=False, noise_std=1.0, white_background=False,
                 blender=True, pdf_padding=True, gaussian_smooth_factor=1.7,
                 dist_reg_coeficient=1.0 / 32, ray_shape="cone", chunksize=16384,
                 loss_coeficients=(1.0, 1.0), dp_coeficient=0.1):
        self.model = model
        self.near = near
        self.far = far
        self.num_coarse = num_coarse
        self.num_fine = num_fine
        self.perturb = perturb
        self.lindisp = lindisp
        self.noise_std = noise_std
        self.white_background = white_background
        self.blender = blender            # dataset.type == 'blender' or basedir endswith 'segmented'
        self.pdf_padding = pdf_padding
        self.gaussian_smooth_factor = gaussian_smooth_factor
        self.dist_reg_coeficient = dist_reg_coeficient
        self.ray_shape = ray_shape
        self.chunksize = chunksize
        self.loss_coeficients = tuple(loss_coeficients)
        self.dp_coeficient = dp_coeficient


# --------------------------------------------------------------------------------------------
# samplers
# --------------------------------------------------------------------------------------------
def sample_first_cycle(near, far, num_coarse, lindisp=False, t_rand=None):
    """samplers.py:30-62.  near/far are [N,1]; ``t_rand`` [N,num_coarse+1] is the stratified
    jitter (``None`` = cfg.nerf[mode].perturb False)."""
    s = torch.linspace(0.0, 1.0, num_coarse + 1, dtype=near.dtype)
    if not lindisp:
        t = near * (1.0 - s) + far * s                                   # samplers.py:40
    else:
        t = 1.0 / (1.0 / near * (1.0 - s) + 1.0 / far * s)               # samplers.py:42
    if t_rand is not None:                                               # samplers.py:52-60
        mids = 0.5 * (t[..., 1:] + t[..., :-1])
        upper = torch.cat((mids, t[..., -1:]), -1)
        lower = torch.cat((t[..., :1], mids), -1)
        t = lower + (upper - lower) * t_rand
        t[:, 0] = near.squeeze(-1)
        t[:, -1] = far.squeeze(-1)
    return t


def resampling_cdf(weights, pdf_padding):
    """Weight smoothing + CDF shared by both resamplers, samplers.py:69-91 / 130-152."""
    pad = torch.cat([weights[..., :1], weights, weights[..., -1:]], -1)
    if pdf_padding:
        mx = torch.maximum(pad[..., :-1], pad[..., 1:])
        w = 0.5 * (mx[..., :-1] + mx[..., 1:]) + 0.01
    else:
        w = 0.8 * weights + 0.1 * pad[..., :-2] + 0.1 * pad[..., 2:] + 0.01
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.minimum(torch.tensor(1.0), torch.cumsum(pdf[..., :-1], -1))
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf, torch.ones_like(cdf[..., :1])], -1)


def mip_u(n_rays, num_samples, rand=None, dtype=torch.float32):
    """samplers.py:94-104.  ``rand`` [N,n] uniform draws or None for det=True."""
    if rand is None:
        return torch.linspace(0.0, 1.0, num_samples, dtype=dtype).expand(n_rays, num_samples)
    s = 1 / num_samples
    u = (torch.arange(num_samples) * s).expand(n_rays, num_samples)
    u = u + rand / ((1 / s) + 1e-5)
    return torch.minimum(u, torch.tensor(0.9999))


def dd_u(n_rays, num_samples, rand=None, dtype=torch.float32):
    """samplers.py:155-171."""
    if rand is None:
        return torch.linspace(0.0, 0.9999, num_samples, dtype=dtype).expand(n_rays, num_samples)
    s = 1 / (num_samples - 1)
    u = (torch.arange(num_samples) * s).expand(n_rays, num_samples)
    u = u + rand / (num_samples + 1e-5)
    u = torch.minimum(u, torch.tensor(0.9999))
    return torch.maximum(u, torch.tensor(0.0))


def _first_equal(sorted_vals, j):
    """torch.max returns the FIRST index of the maximum: among a run of equal values ending at
    j pick the run's first index (samplers.py:178, dd_utils.py:48)."""
    v = torch.gather(sorted_vals, -1, j)
    return torch.searchsorted(sorted_vals.contiguous(), v.contiguous(), right=False)


def find_interval(cdf, u):
    """Interval search of samplers.py:106-116: j = last index with cdf[j] <= u, j1 = first
    index with cdf > u (or the last index)."""
    j = torch.searchsorted(cdf.contiguous(), u.contiguous(), right=True) - 1
    j = j.clamp(min=0)
    j1 = (j + 1).clamp(max=cdf.shape[-1] - 1)
    return j, j1


def sample_pdf(bins, weights, num_samples, pdf_padding, rand=None):
    """mip-NeRF resampler, samplers.py:64-121.  Returns (samples [N,n], j [N,n] int64)."""
    cdf = resampling_cdf(weights, pdf_padding)
    u = mip_u(cdf.shape[0], num_samples, rand, weights.dtype)
    j, j1 = find_interval(cdf, u)
    b0, b1 = torch.gather(bins, -1, j), torch.gather(bins, -1, j1)
    c0, c1 = torch.gather(cdf, -1, j), torch.gather(cdf, -1, j1)
    t = torch.clip(torch.nan_to_num((u - c0) / (c1 - c0), 0), 0, 1)     # samplers.py:118
    return b0 + t * (b1 - b0), j


def sample_pdf_with_mu_sigma(bins, weights, mus, sigmas, part_inside, left_tail, num_samples,
                             pdf_padding, near_cfg, far_cfg, rand=None):
    """DDNeRF resampler, samplers.py:124-215.  Returns (sorted samples [N,n], ind [N,n])."""
    cdf = resampling_cdf(weights, pdf_padding)
    u = dd_u(cdf.shape[0], num_samples, rand, weights.dtype)
    if bins.shape[1] == 2:                                               # samplers.py:185-190
        z = u * part_inside + left_tail
        new_mus, new_sigmas = mus, sigmas
        b0, b1 = bins[:, 0:1], bins[:, 1:2]
        ind = torch.zeros_like(u, dtype=torch.int64)
    else:
        j, j1 = find_interval(cdf, u)
        b0, b1 = torch.gather(bins, -1, j), torch.gather(bins, -1, j1)
        c0, c1 = torch.gather(cdf, -1, j), torch.gather(cdf, -1, j1)
        ind = _first_equal(bins, j)                                      # argmax over bins values
        pin = torch.gather(part_inside, -1, ind)
        lt = torch.gather(left_tail, -1, ind)
        z = ((u - c0) / (c1 - c0)) * pin + lt                            # samplers.py:198
        z = torch.minimum(z, torch.tensor(0.999))
        new_mus = torch.gather(mus, -1, ind)
        new_sigmas = torch.gather(sigmas, -1, ind)
    z = SQRT2 * torch.erfinv(2 * z - 1)                                  # math_utils.py:202-208
    t = torch.clip(z * new_sigmas + new_mus, 0, 0.99999)
    samples = b0 + t * (b1 - b0)
    samples = samples.clone()
    samples[:, -1] = far_cfg                                             # samplers.py:210-211
    samples[:, 0] = near_cfg
    return torch.sort(samples, dim=1)[0], ind


# --------------------------------------------------------------------------------------------
# encoding
# --------------------------------------------------------------------------------------------
def cast_rays(t_vals, origins, directions, radii, ray_shape="cone"):
    """math_utils.py:7-110 (diag=True)."""
    t0, t1 = t_vals[..., :-1], t_vals[..., 1:]
    d = directions
    if ray_shape == "cone":                                              # math_utils.py:76-82
        mu = (t0 + t1) / 2
        hw = (t1 - t0) / 2
        t_mean = mu + (2 * mu * hw ** 2) / (3 * mu ** 2 + hw ** 2)
        t_var = (hw ** 2) / 3 - (4 / 15) * ((hw ** 4 * (12 * mu ** 2 - hw ** 2)) /
                                            (3 * mu ** 2 + hw ** 2) ** 2)
        r_var = radii ** 2 * ((mu ** 2) / 4 + (5 / 12) * hw ** 2 - 4 / 15 *
                              (hw ** 4) / (3 * mu ** 2 + hw ** 2))
    elif ray_shape == "cylinder":                                        # math_utils.py:107-109
        t_mean = (t0 + t1) / 2
        r_var = radii ** 2 / 4
        t_var = (t1 - t0) ** 2 / 12
    else:
        raise AssertionError(ray_shape)
    mean = d[..., None, :] * t_mean[..., None]                           # math_utils.py:36
    d_mag_sq = torch.maximum(torch.tensor(1e-10), torch.sum(d ** 2, -1, keepdim=True))
    d_outer_diag = d ** 2
    null_outer_diag = 1 - d_outer_diag / d_mag_sq
    cov = t_var[..., None] * d_outer_diag[..., None, :] + r_var[..., None] * null_outer_diag[..., None, :]
    return mean + origins[..., None, :], cov


def integrated_pos_enc(means, covs, max_deg=16, min_deg=0):
    """math_utils.py:112-166 -> [N,S,96] = [sin block | cos block], degree-major, xyz-minor."""
    scales = torch.tensor([2 ** i for i in range(min_deg, max_deg)])
    shape = list(means.shape[:-1]) + [-1]
    y = torch.reshape(means[..., None, :] * scales[:, None], shape)
    y_var = torch.reshape(covs[..., None, :] * scales[:, None] ** 2, shape)
    x = torch.cat([y, y + 0.5 * torch.tensor(math.pi)], -1)
    x_var = torch.cat([y_var, y_var], -1)
    t = 100 * torch.tensor(math.pi)                                      # safe_trig_helper :155
    safe = torch.where(torch.abs(x) < t, x, x % t)
    return torch.exp(-0.5 * x_var) * torch.sin(safe)


def positional_encoding(v, num_encoding_functions=4):
    """nerf_helpers.py:127-171 with include_input=True, log_sampling=True -> [...,27]."""
    enc = [v]
    for i in range(num_encoding_functions):
        freq = 2.0 ** i
        enc.append(torch.sin(v * freq))
        enc.append(torch.cos(v * freq))
    return torch.cat(enc, -1)


def encode_rows(rays, t_vals, ray_shape="cone"):
    """run_network's feature build, models.py:117-133: rays [N,12] -> [N*S,123]."""
    ro, rd, rr = rays[..., :3], rays[..., 3:6], rays[..., 6].reshape(-1, 1)
    means, covs = cast_rays(t_vals, ro, rd, rr, ray_shape)
    enc = integrated_pos_enc(means, covs)
    S = enc.shape[1]
    dirs = positional_encoding(rays[..., -3:])
    dirs = dirs[:, None, :].expand(dirs.shape[0], S, dirs.shape[-1])
    return torch.cat((enc.reshape(-1, 96), dirs.reshape(-1, 27)), -1)


# --------------------------------------------------------------------------------------------
# MLP (base_architectures.py:40-61 / 103-126).  ``params`` = state_dict-style mapping.
# --------------------------------------------------------------------------------------------
def mlp_forward(params, x):
    xyz, dirs = x[..., :96], x[..., 96:]
    h = F.relu(F.linear(xyz, params["layers_xyz.0.weight"], params["layers_xyz.0.bias"]))
    for i in range(1, 8):
        inp = torch.cat((xyz, h), -1) if i == 5 else h
        h = F.relu(F.linear(inp, params[f"layers_xyz.{i}.weight"], params[f"layers_xyz.{i}.bias"]))
    feat = F.linear(h, params["fc_feat.weight"], params["fc_feat.bias"])
    alpha = F.linear(feat, params["fc_alpha.weight"], params["fc_alpha.bias"])
    hd = F.relu(F.linear(torch.cat((feat, dirs), -1), params["layers_dir.0.weight"],
                         params["layers_dir.0.bias"]))
    rgb = F.linear(hd, params["fc_rgb.weight"], params["fc_rgb.bias"])
    out = [rgb, alpha]
    if "fc_mu_sigma.weight" in params:
        out.append(F.linear(hd, params["fc_mu_sigma.weight"], params["fc_mu_sigma.bias"]))
    return torch.cat(out, -1)


def init_mlp_params(depth_head, seed):
    """Default nn.Linear initialisation under a seed, same module construction order as
    base_architectures.py:24-38 / 85-100 so a seeded reference model has identical weights."""
    g = torch.Generator().manual_seed(seed)
    shapes = [("layers_xyz.0", 256, 96)]
    for i in range(1, 8):
        shapes.append((f"layers_xyz.{i}", 256, 352 if i == 5 else 256))
    shapes += [("fc_feat", 256, 256), ("fc_alpha", 1, 256), ("layers_dir.0", 128, 283),
               ("fc_rgb", 3, 128)]
    if depth_head:
        shapes.append(("fc_mu_sigma", 2, 128))
    params = {}
    for name, o, i in shapes:
        bound = 1.0 / math.sqrt(i)
        params[name + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        params[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * bound
    return params


def run_network(params, rays, t_vals, ray_shape="cone", chunksize=None):
    """models.py:117-142."""
    x = encode_rows(rays, t_vals, ray_shape)
    if chunksize is None:
        out = mlp_forward(params, x)
    else:
        out = torch.cat([mlp_forward(params, x[i:i + chunksize]) for i in range(0, x.shape[0], chunksize)], 0)
    return out.reshape(rays.shape[0], t_vals.shape[1] - 1, out.shape[-1])


# --------------------------------------------------------------------------------------------
# volume rendering (volume_rendering_utils.py:6-84, nerf_helpers.py:43-64)
# --------------------------------------------------------------------------------------------
def cumprod_exclusive(x):
    c = torch.cumprod(x, -1)
    c = torch.roll(c, 1, -1)
    c[..., 0] = 1.0
    return c


def volume_render(raw, t_vals, rd, noise=None, white_background=False, blender=True, mus=None):
    """``noise`` = randn * std already scaled ([N,S]) or None.  Returns the reference's
    7-tuple (rgb_map, disp, acc, weights, depth, corrected_disp | None, rgb)."""
    mids = (t_vals[..., 1:] + t_vals[..., :-1]) / 2
    dists = t_vals[..., 1:] - t_vals[..., :-1]
    delta = dists * rd[..., None, :].norm(p=2, dim=-1)
    rgb = torch.sigmoid(raw[..., :3]) * (1 + 2 * 0.001) - 0.001
    density = raw[..., 3] + (noise if noise is not None else 0.0)
    sigma_a = F.softplus(density - 1)
    alpha = 1.0 - torch.exp(-sigma_a * delta)
    weights = alpha * cumprod_exclusive(1.0 - alpha + 1e-10)
    rgb_map = (weights[..., None] * rgb).sum(-2)
    if blender:                                                          # :50-58
        eps_mask = torch.zeros_like(weights)
        eps_mask[:, -1] += 1e-10
        weights = weights + eps_mask.detach()
        pdf = weights / (weights.sum(1).reshape(-1, 1))
    else:
        pdf = weights
    depth = (pdf * mids).sum(-1)
    acc = weights.sum(-1)
    disp = 1.0 / torch.max(1e-10 * torch.ones_like(depth), depth / acc)
    if white_background:
        rgb_map = rgb_map + (1.0 - acc[..., None])
    cdisp = None
    if mus is not None:                                                  # :76-83
        cdepth = (pdf * (t_vals[..., :-1] + mus * dists)).sum(1)
        cdisp = 1.0 / torch.max(1e-10 * torch.ones_like(cdepth), cdepth / acc)
        depth = cdepth
    return rgb_map, disp, acc, weights, depth, cdisp, rgb


# --------------------------------------------------------------------------------------------
# depth-distribution loss (dd_utils.py:6-78)
# --------------------------------------------------------------------------------------------
def normal_cdf(x):
    """math_utils.py:193-200 (exact, despite the reference's name 'approximate_cdf')."""
    return 0.5 * (1 + torch.erf(x / torch.sqrt(torch.tensor(2.0))))


def estimate_dp_loss(t1, t0, pdf_1, pdf_0, mus_0, sigmas_0, left_tails_0, part_inside_0, blender):
    if blender:                                                          # dd_utils.py:12-28
        rel = pdf_1.sum(1) > 1e-10
        if rel.sum() == 0:
            return rel.sum().detach()
        pdf_0, pdf_1, mus_0, sigmas_0 = pdf_0[rel], pdf_1[rel], mus_0[rel], sigmas_0[rel]
        part_inside_0, t1, t0 = part_inside_0[rel], t1[rel], t0[rel]
        # NB the reference does NOT filter left_tails_0 (dd_utils.py:22-28); when every row is
        # relevant (always, with the 1e-10 on the last weight) this is the identity.
    eps = 1e-12
    pdf_0 = (pdf_0 + eps) / torch.sum(pdf_0 + eps, -1, keepdim=True)
    pdf_1 = (pdf_1 + eps) / torch.sum(pdf_1 + eps, -1, keepdim=True)
    width = t0[:, 1:] - t0[:, :-1]
    mus_ray = t0[:, :-1] + mus_0 * width
    sig_ray = sigmas_0 * width
    cdf = torch.minimum(torch.tensor(1.0), torch.cumsum(pdf_0[..., :-1], -1))
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf, torch.ones_like(cdf[..., :1])], -1)
    # cell of each fine edge: mask = t1 > t0 (strict), index = first maximum of the masked cdf
    cnt = torch.searchsorted(t0.contiguous(), t1.contiguous(), right=False)
    j = (cnt - 1).clamp(min=0)
    idx = _first_equal(cdf.detach(), j)
    est = torch.gather(cdf, -1, idx)
    mus = torch.gather(mus_ray, -1, idx)
    sig = torch.gather(sig_ray, -1, idx)
    pin = torch.gather(part_inside_0, -1, idx)
    lt = torch.gather(left_tails_0, -1, idx)
    pdf = torch.gather(pdf_0, -1, idx)
    est = est + ((normal_cdf((t1 - mus) / sig) - lt) / pin) * pdf
    est = torch.where(est > 1, torch.ones_like(est), est)               # dd_utils.py:66
    q = est[:, 1:] - est[:, :-1]
    q = torch.where(q < 0, torch.zeros_like(q), q)                      # dd_utils.py:70
    q = (q + eps) / torch.sum(q + eps, -1, keepdim=True)
    with warnings.catch_warnings():                                      # dd_utils.py:76 (default 'mean')
        warnings.simplefilter("ignore")
        return F.kl_div(q.log(), pdf_1.detach(), reduction="mean")


# --------------------------------------------------------------------------------------------
# frame post-processing (row f4)
# --------------------------------------------------------------------------------------------
def cast_to_disparity_image(disp):
    """validation_utils/visualization.py:11-17 -> uint8 [H,W]."""
    img = (disp - disp.min()) / (disp.max() - disp.min())
    img = img.clamp(0, 1) * 255
    return img.detach().cpu().numpy().astype("uint8")


def cast_to_image(rgb):
    """validation_utils/visualization.py:20-27 -> uint8 [H,W,3]: ToPILImage on a float tensor is mul(255).byte().
    The clamp states what the renderer's range (-0.001 .. 1.001, volume_rendering_utils.py:25-27) needs; the
    reference leaves the two out-of-range ends to the float->uint8 cast."""
    return rgb.detach().cpu().mul(255).clamp(0, 255).byte().numpy()


def video_frame(rgb, disp8):
    """render_video.py:96-101 -> uint8 [H, 2W, 3], BGR: the colour image next to the grey disparity image."""
    import numpy as np
    rgb8 = cast_to_image(rgb)
    d3 = np.repeat(disp8[..., None], 3, axis=-1)
    return np.concatenate([rgb8[..., ::-1], d3], axis=1)


# --------------------------------------------------------------------------------------------
# model orchestration (models.py:40-162, 207-322)
# --------------------------------------------------------------------------------------------
def pack_rays(ro, rd, rad, near, far):
    """get_rays_batches, models.py:144-158 -> [N,12]."""
    ro, rd, rad = ro.reshape(-1, 3), rd.reshape(-1, 3), rad.reshape(-1, 1)
    viewdirs = rd / rd.norm(p=2, dim=-1).unsqueeze(-1)
    nr = near * torch.ones_like(rd[..., :1])
    fr = far * torch.ones_like(rd[..., :1])
    return torch.cat((ro, rd, rad, nr, fr, viewdirs), -1)


def predict_mip(cfg, params, rays, rnd):
    """GeneralMipNerfModel.predict, models.py:75-114.  ``rnd`` = dict of injected randoms:
    t_rand [N,Nc+1] | None, u_rand [N,Nf+1] | None, noise0/noise1 [N,S] unit normal | None."""
    near, far = rays[:, 7:8], rays[:, 8:9]
    rd = rays[:, 3:6]
    out = {}
    weights = t = None
    for i in range(2):
        if i == 0:
            t = sample_first_cycle(near, far, cfg.num_coarse, cfg.lindisp, rnd.get("t_rand"))
        else:
            t, _ = sample_pdf(t, weights, cfg.num_fine + 1, cfg.pdf_padding, rnd.get("u_rand"))
            t = t.detach()
        raw = run_network(params, rays, t, cfg.ray_shape)
        nz = rnd.get(f"noise{i}")
        nz = nz * cfg.noise_std if (nz is not None and cfg.noise_std > 0) else None
        rgb, disp, acc, weights, depth, _, _ = volume_render(raw, t, rd, nz, cfg.white_background, cfg.blender)
        out[i] = {"rgb": rgb, "disp": disp, "acc": acc, "weights": weights, "depth": depth, "t_vals": t}
    return out


def predict_dd(cfg, params_coarse, params_fine, rays, rnd):
    """DDNerfModel.predict, models.py:207-322."""
    near, far = rays[:, 7:8], rays[:, 8:9]
    rd = rays[:, 3:6]
    out = {}
    t = sample_first_cycle(near, far, cfg.num_coarse, cfg.lindisp, rnd.get("t_rand"))
    raw = run_network(params_coarse, rays, t, cfg.ray_shape)
    raw_mus, raw_sig = raw[:, :, -2], raw[:, :, -1]
    mus = torch.sigmoid(raw_mus)
    sigmas = torch.sigmoid(raw_sig) + 0.001
    sig_loss = (torch.abs(raw_sig) ** 2).sum() / raw_sig.shape[0]
    mus_loss = (torch.abs(raw_mus) ** 2).sum() / raw_mus.shape[0]
    mus_reg = cfg.dist_reg_coeficient * mus_loss
    sig_reg = cfg.dist_reg_coeficient * sig_loss
    left_tail = normal_cdf((0 - mus) / sigmas)
    part_inside = normal_cdf((1 - mus) / sigmas) - left_tail
    nz = rnd.get("noise0")
    nz = nz * cfg.noise_std if (nz is not None and cfg.noise_std > 0) else None
    rgb, disp, acc, w0, depth, cdisp, _ = volume_render(raw[:, :, :-2], t, rd, nz, cfg.white_background,
                                                        cfg.blender, mus=mus)
    sm_sig = sigmas * cfg.gaussian_smooth_factor
    sm_lt = normal_cdf((0 - mus) / sm_sig)
    sm_pin = normal_cdf((1 - mus) / sm_sig) - sm_lt
    pdf = w0 / torch.sum(w0, -1, keepdim=True)
    sel = pdf > 0.1
    out[0] = {"rgb": rgb, "disp": disp, "acc": acc, "weights": w0, "depth": depth, "mus": mus[sel],
              "sigmas": sigmas[sel], "dp_loss": None, "corrected_disp_map": cdisp,
              "smoothed_sigmas": sm_sig[sel], "mus_loss": mus_loss.unsqueeze(0),
              "sig_loss": sig_loss.unsqueeze(0), "mus_reg": mus_reg.unsqueeze(0),
              "sig_reg": sig_reg.unsqueeze(0), "t_vals": t, "mus_full": mus, "sigmas_full": sigmas}
    t1, _ = sample_pdf_with_mu_sigma(t, w0, mus, sm_sig, sm_pin, sm_lt, cfg.num_fine + 1, cfg.pdf_padding,
                                     cfg.near, cfg.far, rnd.get("u_rand"))
    t1 = t1.detach()
    raw1 = run_network(params_fine, rays, t1, cfg.ray_shape)
    nz = rnd.get("noise1")
    nz = nz * cfg.noise_std if (nz is not None and cfg.noise_std > 0) else None
    rgb1, disp1, acc1, w1, depth1, _, _ = volume_render(raw1, t1, rd, nz, cfg.white_background, cfg.blender)
    dp = estimate_dp_loss(t1.detach(), t.detach(), w1.detach(), w0, mus, sigmas, left_tail.detach(),
                          part_inside.detach(), cfg.blender) * (t1.shape[1] - 1)
    dp = (dp + mus_reg + sig_reg).unsqueeze(0)
    out[1] = {"rgb": rgb1, "disp": disp1, "acc": acc1, "weights": w1, "depth": depth1, "mus": mus[sel],
              "sigmas": sigmas[sel], "dp_loss": dp, "corrected_disp_map": None,
              "smoothed_sigmas": sm_sig[sel], "t_vals": t1}
    return out


def train_loss(cfg, out, target):
    """train_model.py:156-167."""
    loss = 0.0
    for j in range(2):
        loss = loss + cfg.loss_coeficients[j] * F.mse_loss(out[j]["rgb"], target)
    if cfg.model == "DDNerfModel":
        loss = loss + cfg.dp_coeficient * out[1]["dp_loss"].mean()
    return loss


def train_step(cfg, params_coarse, params_fine, rays, target, rnd):
    """One forward + backward of train_model.py:154-170 on one ray chunk.  Returns
    (loss, out, grads_coarse, grads_fine) with grads as name->tensor dicts."""
    pc = {k: v.detach().clone().requires_grad_(True) for k, v in params_coarse.items()}
    if cfg.model == "DDNerfModel":
        pf = {k: v.detach().clone().requires_grad_(True) for k, v in params_fine.items()}
        out = predict_dd(cfg, pc, pf, rays, rnd)
    else:
        pf = None
        out = predict_mip(cfg, pc, rays, rnd)
    loss = train_loss(cfg, out, target)
    loss.backward()
    gc = {k: v.grad for k, v in pc.items()}
    gf = {k: v.grad for k, v in pf.items()} if pf is not None else None
    return loss.detach(), out, gc, gf
