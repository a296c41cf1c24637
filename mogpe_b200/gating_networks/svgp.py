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
                             eps_mc=None, eps_u=None, rng=None):
        """[N, K]: Bernoulli [p, 1 - p] analytic; Softmax Monte-Carlo (gating_networks/svgp.py:69-105)."""
        if num_inducing_samples is not None:
            # [S, N, K]: gating functions given inducing samples (CUDA), likelihood expectation per sample
            h_mean, h_var = self.predict_f(Xnew, num_inducing_samples, full_cov=full_cov, eps_u=eps_u)
            return self.predict_mixing_probs_given_h(h_mean, np.broadcast_to(h_var, h_mean.shape), eps_mc=eps_mc, rng=rng)
        return self._require_model()._predict_mixing_probs(Xnew, eps_mc)

    def predict_mixing_probs_given_h(self, h_mean, h_var=None, eps_mc=None, rng=None):
        """gating_networks/svgp.py:85-105: Pr(alpha = k | h) (h_var None) or its expectation under N(h_mean, h_var); K = 2
        returns [p, 1 - p].  Elementwise on values the caller holds, so it runs on the host (the bounds evaluate the same maps
        inside the fused likelihood kernel)."""
        lik = self.likelihood
        if h_var is None:
            p = lik.conditional_mean(h_mean)
        elif isinstance(lik, Softmax):
            p = lik.predict_mean_and_var(h_mean, h_var, eps=eps_mc, rng=rng)[0]
        else:
            p = lik.predict_mean_and_var(h_mean, h_var)[0]
        if self.num_gating_functions == 1:
            p = np.concatenate([p, 1.0 - p], -1)
        return p
