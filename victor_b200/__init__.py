"""victor-b200: B200-native drop-in for the likelihood hot path of seshnadathur/victor.

Same names as ``victor/__init__.py:3-9`` for the classes on the path.
"""
from .model import CCFModel
from .fit import CCFFit
from .cosmology import BackgroundCosmology
from . import utils
from .utils import InputError
from ._version import __version__

__all__ = ["CCFModel", "CCFFit", "BackgroundCosmology", "InputError", "utils", "__version__"]
