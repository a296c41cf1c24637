"""Keras-flavour API (mirrors mogpe/keras/**): SVGPExpert, SVGPGatingNetwork, MixtureOfSVGPExperts."""
from .experts import SVGPExpert  # noqa: F401
from .gating_networks import SVGPGatingNetwork  # noqa: F401
from .mixture_of_experts import MixtureOfSVGPExperts  # noqa: F401

EXPERT_OBJECTS = {"SVGPExpert": SVGPExpert}
GATING_NETWORK_OBJECTS = {"SVGPGatingNetwork": SVGPGatingNetwork}
