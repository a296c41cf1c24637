"""mogpe/keras/experts.py: SVGPExpert with the reference's constructor and methods; arithmetic on the GPU."""
from typing import Optional

import numpy as np

from .._model_core import Normal
from ..gpflow_lite import SVGP, Module
from .gps import predict_f_given_inducing_samples


class ExpertBase(Module):
    """Interface for an individual expert (mogpe/keras/experts.py:24-35)."""

    def predict_dist(self, Xnew, **kwargs):
        raise NotImplementedError


class SVGPExpert(ExpertBase):
    """Sparse Variational Gaussian Process Expert (mogpe/keras/experts.py:38-250)."""

    def __init__(self, kernel, likelihood, inducing_variable, mean_function=None, num_latent_gps: int = 1,
                 q_diag: bool = False, q_mu=None, q_sqrt=None, whiten: bool = True, name: Optional[str] = "SVGPExpert"):
        self._gp = SVGP(kernel=kernel, likelihood=likelihood, inducing_variable=inducing_variable,
                        mean_function=mean_function, num_latent_gps=num_latent_gps, q_diag=q_diag, q_mu=q_mu,
                        q_sqrt=q_sqrt, whiten=whiten, num_data=None)
        self._gp._owner = self
        self.name = name
        self._parent_model = None

    # the model sets this when the expert is placed in an experts_list
    @property
    def _model(self):
        return self._parent_model

    @property
    def gp(self):
        return self._gp

    def _require_model(self):
        if self._parent_model is None:
            raise RuntimeError("attach the expert to a MixtureOfSVGPExperts before calling predict methods "
                               "(the CUDA engine is built per model)")
        return self._parent_model

    def __call__(self, Xnew, num_inducing_samples: Optional[int] = 1, training: Optional[bool] = False):
        return self.call(Xnew, num_inducing_samples, training)

    def call(self, Xnew, num_inducing_samples: Optional[int] = 1, training: Optional[bool] = False):
        if training:
            return self.predict_dist_given_inducing_samples(Xnew, num_inducing_samples=num_inducing_samples)
        return self.predict_dist(Xnew)

    def predict_f(self, Xnew, num_inducing_samples: Optional[int] = None, full_cov: Optional[bool] = False, eps_u=None):
        """(mean, var): [N, F] each, or [S, N, F] with num_inducing_samples (experts.py:178-192)."""
        if full_cov:
            raise NotImplementedError("full_cov=True is not on the accelerated path")
        if num_inducing_samples:
            return predict_f_given_inducing_samples(Xnew, self.gp, num_inducing_samples, full_cov, eps_u)
        return self._require_model()._predict_f(self, Xnew)

    def predict_y(self, Xnew, num_inducing_samples: int = None, full_cov: bool = False, eps_u=None):
        """Gaussian likelihood push-through: (f_mean, f_var + noise variance) (experts.py:146-158)."""
        f_mean, f_var = self.predict_f(Xnew, num_inducing_samples=num_inducing_samples, full_cov=full_cov, eps_u=eps_u)
        noise = np.broadcast_to(np.asarray(self.gp.likelihood.variance.numpy()).reshape(-1), (self.gp.num_latent_gps,))
        return f_mean, f_var + noise

    def predict_dist_given_y(self, y_mean, y_var):
        """MultivariateNormalDiag(loc=y_mean, scale_diag=y_var): the reference passes the VARIANCE as the scale
        (experts.py:167); reproduced."""
        return Normal(y_mean, y_var, event_ndims=1)

    def predict_dist(self, Xnew, full_cov: bool = False):
        y_mean, y_var = self.predict_y(Xnew, full_cov=full_cov)
        return self.predict_dist_given_y(y_mean, y_var)

    def predict_dist_given_inducing_samples(self, Xnew, num_inducing_samples: int = None, full_cov: bool = False,
                                            eps_u=None):
        y_mean, y_var = self.predict_y(Xnew, num_inducing_samples=num_inducing_samples, full_cov=full_cov, eps_u=eps_u)
        return self.predict_dist_given_y(y_mean, y_var)

    def prior_kl(self):
        """KL[q(u) || p(u)] of this expert (experts.py:205-207)."""
        return self._require_model()._prior_kl(self)

    def get_config(self):
        k = self.gp.kernel
        return {"num_latent_gps": self.gp.num_latent_gps, "q_diag": self.gp.q_diag, "q_mu": self.gp.q_mu.numpy(),
                "q_sqrt": self.gp.q_sqrt.numpy(), "whiten": self.gp.whiten, "kernel": type(k).__name__,
                "inducing_variable": self.gp.Z.numpy()}
