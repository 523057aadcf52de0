"""Latent optimal-transport toolbox (same star-exports as reference ot/__init__.py:1-4, minus the Lightning
callback, which stays in the reference and calls into these classes)."""
from .matrix_utils import *  # noqa: F401,F403
from .w2_utils import *  # noqa: F401,F403
from .distribution_models import *  # noqa: F401,F403
from .transport import *  # noqa: F401,F403
