"""Mirror of the reference's ``general_utils/volume_rendering_utils.py``: one fused compositing
kernel (csrc/composite.cu) behind the reference's signature."""
import torch

from .. import ops


def _is_blender(cfg):
    """volume_rendering_utils.py:50-51."""
    if cfg is None:
        return False
    return cfg.dataset.type.lower() == "blender" or str(cfg.dataset.basedir).endswith("segmented")


def volume_render_radiance_field(radiance_field, depth_values, ray_directions, radiance_field_noise_std=0.0,
                                 white_background=False, mus=None, cfg=None, noise=None, want_rgb=True):
    """volume_rendering_utils.py:6-84.  Returns the reference's 7-tuple
    (rgb_map, disp_map, acc_map, weights, depth_map, corrected_disp_map | None, rgb).

    Extensions (keyword-only in spirit, defaults reproduce the reference): ``noise`` injects the
    unit-normal draw the reference makes at line 31 (same shape ``radiance_field[..., 3]``);
    ``want_rgb=False`` skips materialising the per-sample colours nobody reads on the model path.
    """
    if radiance_field_noise_std > 0.0 and noise is None:
        noise = torch.randn(radiance_field[..., 3].shape, dtype=radiance_field.dtype, device=radiance_field.device)
    if radiance_field_noise_std <= 0.0:
        noise = None
    return ops.composite(radiance_field[..., :4], depth_values, ray_directions, noise, radiance_field_noise_std, mus,
                         white_background, _is_blender(cfg), want_rgb)
