"""mogpe/keras/gps.py: conditional given samples of the inducing outputs."""
from typing import Optional

import numpy as np


def predict_f_given_inducing_samples(Xnew, svgp, num_inducing_samples: int = 1, full_cov: Optional[bool] = False,
                                     eps_u=None):
    """Draw u_s ~ q(u) and evaluate the whitened conditional per sample with q_sqrt=None
    (reference: mogpe/keras/gps.py:14-50) -> mean [S, N, L], var [S, N, L].

    `svgp` must belong to a component attached to a MixtureOfSVGPExperts (the CUDA engine is per model).
    eps_u [S, L, M]: optional explicit standard-normal draws (the reference's are TF-internal)."""
    if full_cov:
        raise NotImplementedError("full_cov=True is not on the accelerated path")
    owner = getattr(svgp, "_owner", None)
    if owner is None or owner._model is None:
        raise RuntimeError("attach the component to a MixtureOfSVGPExperts before calling predict methods")
    return owner._model._predict_f_inducing(owner, Xnew, num_inducing_samples, eps_u)
