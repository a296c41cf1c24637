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
                             eps_mc=None, eps_u=None, rng=None):
        """Pr(alpha = k | Xnew) [N, K] (gating_networks.py:173-186): Bernoulli analytic, Softmax Monte-Carlo."""
        if num_inducing_samples:
            # [S, N, K]: gating functions given inducing samples (CUDA), likelihood expectation per sample (gating_networks.py:173-186)
            h_mean, h_var = self.predict_h(Xnew, num_inducing_samples=num_inducing_samples, eps_u=eps_u)
            return self.predict_mixing_probs_given_h(h_mean, np.broadcast_to(h_var, h_mean.shape), eps_mc=eps_mc, rng=rng)
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

    def predict_mixing_probs_given_h(self, h_mean, h_var=None, eps_mc=None, rng=None):
        """gating_networks.py:228-249: Pr(alpha = k | h = h_mean), or its expectation under N(h_mean, h_var) when h_var is
        given (Bernoulli: analytic probit; Softmax: gpflow's Monte-Carlo mean, eps_mc [R, ..., K] for explicit draws).
        Elementwise on gating-function values the caller holds (for K = 2 they are the [h, -h] pairs of predict_h), so it
        runs on the host; the bounds evaluate the same maps inside the fused likelihood kernel."""
        lik = self.gp.likelihood
        if h_var is None:
            return lik.conditional_mean(h_mean)
        if isinstance(lik, Softmax):
            return lik.predict_mean_and_var(h_mean, h_var, eps=eps_mc, rng=rng)[0]
        return lik.predict_mean_and_var(h_mean, h_var)[0]

    def predict_categorical_dist_given_h_samples(self, Xnew, num_h_samples: Optional[int] = 1, eps_h=None, rng=None):
        """gating_networks.py:145-157: Categorical over the expert indicator for every gating-function sample, batch shape
        [S, N] (the posterior h comes from the CUDA predict path, the samples and the link are elementwise)."""
        h = self.predict_h_samples(Xnew, num_samples=num_h_samples, eps_h=eps_h, rng=rng)
        return Categorical(self.predict_mixing_probs_given_h(h))

    def predict_categorical_dist_given_inducing_samples(self, Xnew, num_inducing_samples: Optional[int] = 1, eps_u=None,
                                                        eps_mc=None, rng=None):
        """gating_networks.py:159-171: h given inducing samples (CUDA: mogpe_predict_f_inducing_samples), then the likelihood
        expectation."""
        h_mean, h_var = self.predict_h(Xnew, num_inducing_samples=num_inducing_samples, eps_u=eps_u)
        return Categorical(self.predict_mixing_probs_given_h(h_mean, np.broadcast_to(h_var, h_mean.shape), eps_mc=eps_mc, rng=rng))

    def prior_kl(self):
        return self._require_model()._prior_kl(self)
