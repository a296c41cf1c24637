"""mogpe_b200: B200-native hot path (minibatch ELBO fwd+bwd, predict) of aidanscannell/mogpe.

Legacy-flavour API at the top level (mirrors ``mogpe``), Keras-flavour API under ``mogpe_b200.keras``
(mirrors ``mogpe.keras``).  All arithmetic runs in the CUDA library behind include/mogpe_b200.h.
"""
__version__ = "0.1.0"
