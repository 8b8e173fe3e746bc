"""Seeded synthetic inputs shared by the golden generator and the tests."""
import torch

from ddnerf_b200.rays import synth_rays  # noqa: F401  (re-export)


def peaked_weights(n_rays, n_samples, seed=0):
    """Weights that exercise the tails / clamps / ties: w = rand^4 plus a Gaussian bump per ray,
    with runs of exact zeros (empty space) so CDFs contain repeated values."""
    g = torch.Generator().manual_seed(seed)
    w = torch.rand(n_rays, n_samples, generator=g) ** 4 * 0.2
    centre = torch.rand(n_rays, 1, generator=g) * n_samples
    width = torch.rand(n_rays, 1, generator=g) * 2 + 0.5
    x = torch.arange(n_samples)[None, :].float()
    w = w + torch.exp(-0.5 * ((x - centre) / width) ** 2)
    zero_from = torch.randint(0, n_samples, (n_rays, 1), generator=g)
    zero_len = torch.randint(0, n_samples // 2, (n_rays, 1), generator=g)
    mask = (x >= zero_from) & (x < zero_from + zero_len)
    w = torch.where(mask, torch.zeros_like(w), w)
    return w / w.sum(-1, keepdim=True).clamp(min=1e-6) * 0.9
