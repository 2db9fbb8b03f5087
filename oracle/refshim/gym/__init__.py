"""Stub of the ``gym`` package: the reference envs only build ``gym.spaces.Box`` objects
(envs/dflex_env.py:18,48-49).  Test infrastructure only."""
from . import spaces  # noqa: F401
