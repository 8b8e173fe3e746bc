from .cfgnode import CfgNode  # noqa: F401
from .nerf_helpers import *  # noqa: F401,F403
from .volume_rendering_utils import *  # noqa: F401,F403
