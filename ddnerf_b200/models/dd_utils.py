"""Mirror of the reference's ``models/dd_utils.py``: the depth-distribution loss as one fused
forward and one fused backward kernel (csrc/dploss.cu)."""
from ..general_utils.math_utils import approximate_cdf  # noqa: F401  (dd_utils.py:2)
from .. import ops


def estimate_dp_loss(t_vals_1, t_vals_0, pdf_1, pdf_0, mus_0, sigmas_0, left_tails_0, part_inside_cells_0, cfg=None):
    """dd_utils.py:6-78.  Gradients flow to pdf_0, mus_0, sigmas_0 (the reference detaches the rest
    at its call site, models.py:287-288).  ``left_tails_0 = part_inside_cells_0 = None`` (this package's fused
    DDNeRF path): the kernels evaluate Phi((0 - mu) / sigma) and Phi((1 - mu) / sigma) per cell themselves."""
    blender = cfg.dataset.type.lower() == "blender"
    lt = None if left_tails_0 is None else left_tails_0.detach()
    pin = None if part_inside_cells_0 is None else part_inside_cells_0.detach()
    return ops.dp_loss(t_vals_1.detach(), t_vals_0.detach(), pdf_1.detach(), pdf_0, mus_0, sigmas_0, lt, pin, blender)
