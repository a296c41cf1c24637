"""mogpe/keras/gating_networks.py: SVGPGatingNetwork with the reference's constructor and methods."""
from typing import Optional

import numpy as np

from .._model_core import Categorical
from ..gpflow_lite import SVGP, Bernoulli, Module, MultioutputKernel, Softmax
from .gps import predict_f_given_inducing_samples


class GatingNetworkBase(Module):
    def __init__(self, num_experts: int, name: Optional[str] = "gating_network"):
        self._num_experts = num_experts
        self.name = name

    def predict_categorical_dist(self, Xnew, **kwargs):
        return Categorical(self.predict_mixing_probs(Xnew, **kwargs))

    def predict_mixing_probs(self, Xnew, **kwargs):
        raise NotImplementedError

    @property
    def num_experts(self):
        return self._num_experts


class SVGPGatingNetwork(GatingNetworkBase):
    """Bernoulli gating (one GP, K = 2) for a single-output kernel, Softmax gating (K GPs) for a
    MultioutputKernel -- inferred from the kernel as the reference does (gating_networks.py:99-106)."""

    def __init__(self, kernel, inducing_variable, mean_function=None, q_diag: bool = False, q_mu=None, q_sqrt=None,
                 whiten: bool = True, name: Optional[str] = "SVGPGatingNetwork"):
        if isinstance(kernel, MultioutputKernel):
            self.num_gating_gps = kernel.num_latent_gps
            num_experts = self.num_gating_gps
            likelihood = Softmax(num_classes=self.num_gating_gps)
        else:
            self.num_gating_gps = 1
            num_experts = 2
            likelihood = Bernoulli()
        self._gp = SVGP(kernel=kernel, likelihood=likelihood, inducing_variable=inducing_variable,
                        mean_function=mean_function, num_latent_gps=self.num_gating_gps, q_diag=q_diag, q_mu=q_mu,
                        q_sqrt=q_sqrt, whiten=whiten, num_data=None)
        self._gp._owner = self
        super().__init__(num_experts=num_experts, name=name)
        self._parent_model = None

    @property
    def _model(self):
        return self._parent_model

    @property
    def gp(self):
        return self._gp

    def _require_model(self):
        if self._parent_model is None:
            raise RuntimeError("attach the gating network to a MixtureOfSVGPExperts before calling predict methods")
        return self._parent_model

    def __call__(self, Xnew, num_samples: Optional[int] = 1, training: Optional[bool] = False):
        if training:
            return self.predict_categorical_dist_given_h_samples(Xnew, num_h_samples=num_samples)
        return self.predict_categorical_dist(Xnew)

    call = __call__

    def predict_f(self, Xnew, full_cov: Optional[bool] = False):
        if full_cov:
            raise NotImplementedError("full_cov=True is not on the accelerated path")
        return self._require_model()._predict_f(self, Xnew)

    def predict_h(self, Xnew, num_inducing_samples: Optional[int] = None, full_cov: Optional[bool] = False, eps_u=None):
        """Gating function posterior; K = 2 returns [h, -h] / [v, v] (gating_networks.py:201-226)."""
        if full_cov:
            raise NotImplementedError("full_cov=True is not on the accelerated path")
        if num_inducing_samples:
            h_mean, h_var = predict_f_given_inducing_samples(Xnew, self.gp, num_inducing_samples, full_cov, eps_u)
        else:
            h_mean, h_var = self.predict_f(Xnew)
        if self.num_gating_gps == 1:
            h_mean = np.concatenate([h_mean, -h_mean], -1)
            h_var = np.concatenate([h_var, h_var], -1)
        return h_mean, h_var

    def predict_mixing_probs(self, Xnew, num_inducing_samples: Optional[int] = None, full_cov: Optional[bool] = False,
                             eps_mc=None):
        """Pr(alpha = k | Xnew) [N, K] (gating_networks.py:173-186): Bernoulli analytic, Softmax Monte-Carlo."""
        if num_inducing_samples:
            raise NotImplementedError("predict_mixing_probs(num_inducing_samples=...) is only used inside the tight "
                                      "bound, which the engine evaluates internally")
        return self._require_model()._predict_mixing_probs(Xnew, eps_mc)

    def predict_categorical_dist(self, Xnew, **kw):
        return Categorical(self.predict_mixing_probs(Xnew, **kw))

    def predict_h_samples(self, Xnew, num_samples: Optional[int] = None, full_cov: Optional[bool] = False, eps_h=None,
                          rng=None):
        """h_s = mean + sqrt(var) eps (gpflow predict_f_samples), K = 2: [h, -h] (gating_networks.py:188-199)."""
        h_mean, h_var = self.predict_f(Xnew)
        S = num_samples or 1
        if eps_h is None:
            eps_h = (rng or np.random.default_rng()).standard_normal((S,) + h_mean.shape)
        h = h_mean[None] + np.sqrt(h_var)[None] * eps_h
        if self.num_gating_gps == 1:
            h = np.concatenate([h, -h], -1)
        return h

    def predict_mixing_probs_given_h(self, h_mean, h_var=None):
        raise NotImplementedError("evaluated inside the fused likelihood kernel; use predict_mixing_probs / the bounds")

    def predict_categorical_dist_given_h_samples(self, Xnew, num_h_samples: Optional[int] = 1):
        raise NotImplementedError("evaluated inside the fused likelihood kernel; use the model's bounds")

    def prior_kl(self):
        return self._require_model()._prior_kl(self)
