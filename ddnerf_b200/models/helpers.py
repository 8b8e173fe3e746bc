from ..general_utils.nerf_helpers import get_minibatches  # noqa: F401  (models/helpers.py duplicates it)
