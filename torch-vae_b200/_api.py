"""Public surface of the B200-native VanillaVAE training path."""

from . import config, data_transformations, engine, ops, optim, parallel, trainer  # noqa: F401
from ._lib import LIB_PATH, load as load_library  # noqa: F401
from .models import VanillaVAE
from .optim import FusedAdamW
from .types_helpers import EncoderOutput, LossOutput, ModelOutput

__all__ = [
    "VanillaVAE",
    "FusedAdamW",
    "optim",
    "EncoderOutput",
    "ModelOutput",
    "LossOutput",
    "ops",
    "engine",
    "parallel",
    "config",
    "data_transformations",
    "trainer",
    "load_library",
    "LIB_PATH",
]
