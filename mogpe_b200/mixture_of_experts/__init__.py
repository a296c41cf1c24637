from .base import MixtureOfExperts  # noqa: F401
from .svgp import MixtureOfSVGPExperts  # noqa: F401
