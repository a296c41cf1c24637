from .model_parsers import MixtureOfSVGPExperts_from_toml  # noqa: F401
