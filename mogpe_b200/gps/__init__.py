from .svgp import SVGPModel  # noqa: F401
from ..gpflow_lite import Gaussian  # noqa: F401  (mogpe/gps/likelihoods.py)
