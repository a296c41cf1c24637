from .base import GatingNetworkBase  # noqa: F401
from .svgp import SVGPGatingNetwork  # noqa: F401
