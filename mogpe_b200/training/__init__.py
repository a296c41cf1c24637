"""Callers of the hot path that the reference ships next to it (mogpe/training): config front-ends only.
The training loops, TensorBoard monitors and checkpoint managers are out of scope."""
from .toml_config_parsers import MixtureOfSVGPExperts_from_toml  # noqa: F401
