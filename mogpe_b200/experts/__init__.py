from .base import ExpertBase, ExpertsBase  # noqa: F401
from .svgp import SVGPExpert, SVGPExperts  # noqa: F401
