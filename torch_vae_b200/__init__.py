"""Importable alias of the ``torch-vae_b200/`` package directory.

The package directory carries the reference repo's name (with a hyphen), which
Python cannot import directly; this shim points the package search path at it.
"""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "torch-vae_b200"))

from ._api import *  # noqa: F401,F403,E402
from ._api import __all__  # noqa: E402
