"""mogpe/gating_networks/svgp.py: legacy SVGPGatingNetwork (likelihood passed in)."""
from typing import Optional

import numpy as np

from ..gpflow_lite import Bernoulli, Softmax
from ..gps import SVGPModel
from .base import GatingNetworkBase


class SVGPGatingNetwork(SVGPModel, GatingNetworkBase):
    def __init__(self, kernel, likelihood, inducing_variable, mean_function=None, num_gating_functions: Optional[int] = 1,
                 q_diag: Optional[bool] = False, q_mu=None, q_sqrt=None, whiten: Optional[bool] = True, num_data=None):
        super().__init__(kernel=kernel, likelihood=likelihood, inducing_variable=inducing_variable,
                         mean_function=mean_function, num_latent_gps=num_gating_functions, q_diag=q_diag, q_mu=q_mu,
                         q_sqrt=q_sqrt, whiten=whiten, num_data=num_data)
        if isinstance(likelihood, Bernoulli):
            self.num_experts = 2
            assert num_gating_functions == 1
        elif isinstance(likelihood, Softmax):
            assert num_gating_functions > 1
            self.num_experts = num_gating_functions
        else:
            raise AttributeError("likelihood should be Bernoulli or Softmax")
        self.num_gating_functions = num_gating_functions

    def predict_fs(self, Xnew, num_inducing_samples: Optional[int] = None, full_cov: Optional[bool] = False, eps_u=None):
        """K = 2: [f, -f], [v, v] (gating_networks/svgp.py:49-67)."""
        f_mean, f_var = self.predict_f(Xnew, num_inducing_samples, full_cov=full_cov, eps_u=eps_u)
        if self.num_gating_functions == 1:
            return np.concatenate([f_mean, -f_mean], -1), np.concatenate([f_var, f_var], -1)
        return f_mean, f_var

    def predict_mixing_probs(self, Xnew, num_inducing_samples: Optional[int] = None, full_cov: Optional[bool] = False,
                             eps_mc=None):
        """[N, K]: Bernoulli [p, 1 - p] analytic; Softmax Monte-Carlo (gating_networks/svgp.py:69-105)."""
        if num_inducing_samples is not None:
            raise NotImplementedError("predict_mixing_probs(num_inducing_samples=...) is only used inside the tight "
                                      "bound, which the engine evaluates internally")
        return self._require_model()._predict_mixing_probs(Xnew, eps_mc)

    def predict_mixing_probs_given_h(self, h_mean, h_var=None):
        raise NotImplementedError("evaluated inside the fused likelihood kernel; use predict_mixing_probs / the bounds")
