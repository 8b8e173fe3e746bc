"""Mirror of the reference's ``models/samplers.py``: same functions and argument meaning, bodies
are single kernel calls (csrc/sampler.cu).

Random draws: the reference draws ``torch.rand`` inside these functions (samplers.py:57,102,165).
The mirrors draw the same shape/dtype on the same device unless the caller injects the tensor
(``t_rand=`` / ``rand=``), which is how the parity tests feed both sides identical numbers.
"""
import torch

from ..general_utils.math_utils import *  # noqa: F401,F403  (samplers.py:1 re-exports math_utils)
from ..general_utils.nerf_helpers import get_minibatches  # noqa: F401
from .. import ops


def get_combined_samples(cfg, near, far, mode):
    """samplers.py:6-27 (off in every shipped config: combined_sampling_method False).  Host-side
    tensor expression, one row broadcast to all rays."""
    nc = cfg.nerf[mode]["num_coarse"]
    t_vals = torch.linspace(0.0, 1.0, nc // 2 + 1, dtype=near.dtype, device=near.device)
    t_uniform = cfg.dataset.near * (1.0 - t_vals) + cfg.dataset.combined_split * t_vals
    min_d, max_d = cfg.dataset.combined_split, far[0]
    d_i = min_d * (1.0 - t_vals) + max_d * t_vals
    t_non = min_d + torch.sort(1 - (torch.log2(d_i - min_d + 1) / torch.log2(max_d - min_d + 1)))[0] * (max_d - min_d)
    return torch.cat((t_uniform, t_non[1:])).expand((near.shape[0], nc + 1))


def sample_first_cycle(cfg, near, far, mode, t_rand=None):
    """samplers.py:30-62."""
    mcfg = getattr(cfg.nerf, mode)
    nc = cfg.nerf[mode]["num_coarse"]
    perturb = bool(cfg.nerf[mode]["perturb"])
    combined = False
    try:
        combined = bool(cfg.dataset.combined_sampling_method)
    except Exception:
        pass
    if combined:
        # rare path, kept as tensor expressions (samplers.py:44-60)
        t_vals = get_combined_samples(cfg, near, far, mode)
        if perturb:
            mids = 0.5 * (t_vals[..., 1:] + t_vals[..., :-1])
            upper = torch.cat((mids, t_vals[..., -1:]), -1)
            lower = torch.cat((t_vals[..., :1], mids), -1)
            if t_rand is None:
                t_rand = torch.rand(t_vals.shape, dtype=near.dtype, device=near.device)
            t_vals = lower + (upper - lower) * t_rand
            t_vals[:, 0] = near.squeeze()
            t_vals[:, -1] = far.squeeze()
        return t_vals
    if perturb and t_rand is None:
        t_rand = torch.rand((near.shape[0], nc + 1), dtype=near.dtype, device=near.device)
    if not perturb:
        t_rand = None
    return ops.sample_first_cycle(near, far, nc, bool(mcfg.lindisp), t_rand)


def sample_pdf(bins, weights, num_samples, cfg, det=True, rand=None):
    """samplers.py:64-121.  Returns a fresh leaf like the reference's nn.Parameter."""
    if not det and rand is None:
        rand = torch.rand(weights.shape[0], num_samples, device=weights.device)
    if det:
        rand = None
    samples = ops.sample_pdf(bins, weights, num_samples, cfg.train_params.pdf_padding, rand)
    return torch.nn.Parameter(samples)


def sample_pdf_with_mu_sigma(bins, weights, mus, sigmas, part_inside_bins, left_tail, num_samples, cfg, det=True,
                             rand=None):
    """samplers.py:124-215."""
    if not det and rand is None:
        rand = torch.rand(weights.shape[0], num_samples, device=weights.device)
    if det:
        rand = None
    samples = ops.sample_pdf_mu_sigma(bins, weights, mus, sigmas, part_inside_bins, left_tail, num_samples,
                                      cfg.train_params.pdf_padding, cfg.dataset.near, cfg.dataset.far, rand)
    return torch.nn.Parameter(samples)
