"""B200-native training path for the reference's VanillaVAE (import it as
``torch_vae_b200``; this directory name mirrors the reference repo's)."""
from ._api import *  # noqa: F401,F403
from ._api import __all__  # noqa: F401
