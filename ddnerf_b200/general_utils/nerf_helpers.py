"""Mirror of the reference's ``general_utils/nerf_helpers.py`` (hot-path subset: lines 8-24,
43-64, 127-180) plus the ray set-up it exports (``get_ray_bundle``, row f1 of SURVEY.md 8f)."""
import math
from typing import Optional

import torch

from .math_utils import *  # noqa: F401,F403  (the reference re-exports math_utils here, nerf_helpers.py:1)
from ..rays import get_ray_bundle  # noqa: F401


def img2mse(img_src, img_tgt):
    return torch.nn.functional.mse_loss(img_src, img_tgt)


def mse2psnr(mse):
    if mse == 0:
        mse = 1e-5
    return -10.0 * math.log10(mse)


def get_minibatches(inputs: torch.Tensor, chunksize: Optional[int] = 1024 * 8):
    """nerf_helpers.py:19-24."""
    return [inputs[i:i + chunksize] for i in range(0, inputs.shape[0], chunksize)]


def cumprod_exclusive(tensor: torch.Tensor) -> torch.Tensor:
    """nerf_helpers.py:43-64.  (The compositing kernel computes this as a lane-group scan; this
    tensor-level version is kept for API compatibility.)"""
    cumprod = torch.cumprod(tensor, -1)
    cumprod = torch.roll(cumprod, 1, -1)
    cumprod[..., 0] = 1.0
    return cumprod


def positional_encoding(tensor, num_encoding_functions=6, include_input=True, log_sampling=True) -> torch.Tensor:
    """nerf_helpers.py:127-171."""
    encoding = [tensor] if include_input else []
    if log_sampling:
        bands = 2.0 ** torch.linspace(0.0, num_encoding_functions - 1, num_encoding_functions, dtype=tensor.dtype,
                                      device=tensor.device)
    else:
        bands = torch.linspace(2.0 ** 0.0, 2.0 ** (num_encoding_functions - 1), num_encoding_functions,
                               dtype=tensor.dtype, device=tensor.device)
    for freq in bands:
        for func in (torch.sin, torch.cos):
            encoding.append(func(tensor * freq))
    return encoding[0] if len(encoding) == 1 else torch.cat(encoding, -1)


def get_embedding_function(num_encoding_functions=6, include_input=True, log_sampling=True):
    return lambda x: positional_encoding(x, num_encoding_functions, include_input, log_sampling)


def meshgrid_xy(tensor1, tensor2):
    """nerf_helpers.py:27-40: np.meshgrid(..., indexing="xy")."""
    ii, jj = torch.meshgrid(tensor1, tensor2, indexing="ij")
    return ii.transpose(-1, -2), jj.transpose(-1, -2)


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """nerf_helpers.py:182-208 (unused by the shipped drivers, kept for API completeness): shift the
    origins to the near plane, then the perspective projection to normalised device coordinates."""
    from ..rays import ndc_project
    return ndc_project(H, W, focal, near, rays_o, rays_d)


def learning_rate_decay(step, lr_init, lr_final, max_steps, lr_delay_steps=0, lr_delay_mult=1):
    """nerf_helpers.py:211-245: log-linear interpolation lr_init -> lr_final with an optional sinusoidal
    warm-up over the first lr_delay_steps (host scalar math of the driver loop)."""
    if lr_delay_steps > 0:
        delay_rate = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0), 1))
    else:
        delay_rate = 1.0
    t = min(max(step / max_steps, 0), 1)
    return delay_rate * math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)
