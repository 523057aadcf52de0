from .fid import *  # noqa: F401,F403
