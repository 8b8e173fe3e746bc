"""Mirror of the reference's ``general_utils/math_utils.py`` (hot-path subset, lines 7-166 and
193-208).

``cast_rays`` / ``integrated_pos_enc`` exist for API compatibility and tests; the model path
never calls them -- it runs the same arithmetic inside the encode / MLP kernels
(csrc/encode.cuh).  They are plain tensor expressions on whatever device their inputs live on.
"""
import math

import torch


def cast_rays(t_vals, origins, directions, radii, ray_shape="cone", diag=True):
    """math_utils.py:7-31.  Returns (means, covs) of shape [..., S, 3] (diagonal covariances)."""
    if not diag:
        raise NotImplementedError("full covariances are never used by the reference (diag=True everywhere)")
    t0, t1 = t_vals[..., :-1], t_vals[..., 1:]
    if ray_shape == "cone":
        means, covs = conical_frustum_to_gaussian(directions, t0, t1, radii, diag)
    elif ray_shape == "cylinder":
        means, covs = cylinder_to_gaussian(directions, t0, t1, radii, diag)
    else:
        assert False
    return means + origins[..., None, :], covs


def lift_gaussian(d, t_mean, t_var, r_var, diag=True):
    """math_utils.py:34-46."""
    mean = d[..., None, :] * t_mean[..., None]
    d_mag_sq = torch.clamp_min(torch.sum(d ** 2, -1, keepdim=True), 1e-10)
    d_outer_diag = d ** 2
    null_outer_diag = 1 - d_outer_diag / d_mag_sq
    return mean, t_var[..., None] * d_outer_diag[..., None, :] + r_var[..., None] * null_outer_diag[..., None, :]


def conical_frustum_to_gaussian(d, t0, t1, base_radius, diag=True, stable=True):
    """math_utils.py:57-88 (stable form)."""
    mu = (t0 + t1) / 2
    hw = (t1 - t0) / 2
    den = 3 * mu ** 2 + hw ** 2
    t_mean = mu + (2 * mu * hw ** 2) / den
    t_var = (hw ** 2) / 3 - (4 / 15) * ((hw ** 4 * (12 * mu ** 2 - hw ** 2)) / den ** 2)
    r_var = base_radius ** 2 * ((mu ** 2) / 4 + (5 / 12) * hw ** 2 - 4 / 15 * (hw ** 4) / den)
    return lift_gaussian(d, t_mean, t_var, r_var, diag)


def cylinder_to_gaussian(d, t0, t1, radius, diag=True):
    """math_utils.py:91-110."""
    return lift_gaussian(d, (t0 + t1) / 2, (t1 - t0) ** 2 / 12, radius ** 2 / 4, diag)


def safe_trig_helper(x, fn, t=100 * math.pi):
    tt = torch.tensor(t, dtype=x.dtype, device=x.device)
    return fn(torch.where(torch.abs(x) < tt, x, x % tt))


def safe_cos(x):
    return safe_trig_helper(x, torch.cos)


def safe_sin(x):
    return safe_trig_helper(x, torch.sin)


def expected_sin(x, x_var):
    """math_utils.py:146-151."""
    y = torch.exp(-0.5 * x_var) * safe_sin(x)
    y_var = torch.clamp_min(0.5 * (1 - torch.exp(-2 * x_var) * safe_cos(2 * x)) - y ** 2, 0)
    return y, y_var


def integrated_pos_enc(x_coord, max_deg=16, min_deg=0, diag=True):
    """math_utils.py:112-144: (means, covs) -> [..., 6*(max_deg-min_deg)]."""
    if not diag:
        raise NotImplementedError("full covariances are never used by the reference (diag=True everywhere)")
    x, x_cov_diag = x_coord
    scales = torch.tensor([2 ** i for i in range(min_deg, max_deg)], device=x.device)
    shape = list(x.shape[:-1]) + [-1]
    y = torch.reshape(x[..., None, :] * scales[:, None], shape)
    y_var = torch.reshape(x_cov_diag[..., None, :] * scales[:, None] ** 2, shape)
    half_pi = torch.tensor(0.5 * math.pi, dtype=x.dtype, device=x.device)
    return expected_sin(torch.cat([y, y + half_pi], -1), torch.cat([y_var] * 2, -1))[0]


def approximate_cdf(x):
    """math_utils.py:193-200: the exact standard-normal CDF."""
    return 0.5 * (1 + torch.erf(x / math.sqrt(2.0)))


def approximate_inverse_cdf(x):
    """math_utils.py:202-208."""
    return math.sqrt(2.0) * torch.erfinv(2 * x - 1)
