"""The ``dflex`` import surface the reference envs / asset loaders / algorithms rely on
(``dflex/__init__.py:8-15`` star-imports sim, render and util): names, not implementation."""
from . import config, model, render, sim, util  # noqa: F401
from .model import *  # noqa: F401,F403
from .model import Mesh, Model, ModelBuilder, State, model_from_articulation  # noqa: F401
from .render import UsdRenderer  # noqa: F401
from .sim import SemiImplicitIntegrator  # noqa: F401
from .util import *  # noqa: F401,F403
from .util import ScopedTimer  # noqa: F401
