"""Optimizer registry (reference optimizers/__init__.py: ``from torch.optim import *`` + RAdam):
``getattr(articulatory_b200.optimizers, config["generator_optimizer_type"])`` resolves the torch
optimizers by name; ``Adam`` inside the fused train step is the flat-buffer ``FusedAdam``."""
from torch.optim import *  # noqa: F401,F403
from torch.optim import RAdam  # noqa: F401  (torch ships the reference's optimizers/radam.py algorithm)

from ..optim import FusedAdam  # noqa: F401
