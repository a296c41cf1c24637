from .base import MixtureOfExpertsBase  # noqa: F401
from .mosvgpe import MixtureOfSVGPExperts, variational_expectation_scale  # noqa: F401
