"""mogpe/experts/svgp.py: legacy SVGPExpert (an SVGPModel) and the SVGPExperts container."""
from typing import List

import numpy as np

from .._model_core import Normal
from ..gps import SVGPModel
from .base import ExpertBase, ExpertsBase


class SVGPExpert(SVGPModel, ExpertBase):
    def __init__(self, kernel, likelihood, inducing_variable, mean_function=None, num_latent_gps: int = 1,
                 q_diag: bool = False, q_mu=None, q_sqrt=None, whiten: bool = True, num_data=None):
        super().__init__(kernel, likelihood, inducing_variable, mean_function, num_latent_gps, q_diag, q_mu, q_sqrt,
                         whiten, num_data)

    def predict_dist(self, Xnew, num_inducing_samples: int = None, full_cov: bool = False, full_output_cov: bool = False,
                     eps_u=None):
        """(mean, var) of y (experts/svgp.py:49-82)."""
        return self.predict_y(Xnew, num_inducing_samples, full_cov, full_output_cov, eps_u)


class SVGPExperts(ExpertsBase):
    """Set of SVGPExpert experts stacked on a trailing axis (experts/svgp.py:145-239)."""

    def __init__(self, experts_list: List[SVGPExpert] = None, name="Experts"):
        super().__init__(experts_list, name=name)
        for expert in experts_list:
            assert isinstance(expert, SVGPExpert)

    def prior_kls(self) -> np.ndarray:
        return np.array([expert.prior_kl() for expert in self.experts_list])

    def predict_dists(self, Xnew, **kwargs) -> Normal:
        """tfd.Normal(mus, sqrt(vars)) with batch shape [..., N, F, K] (experts/svgp.py:171-184)."""
        mus, vs = zip(*[expert.predict_dist(Xnew, **kwargs) for expert in self.experts_list])
        return Normal(np.stack(mus, -1), np.sqrt(np.stack(vs, -1)))

    def predict_fs(self, Xnew, num_inducing_samples: int = None, full_cov=False, full_output_cov=False):
        mus, vs = zip(*[e.predict_f(Xnew, num_inducing_samples, full_cov, full_output_cov) for e in self.experts_list])
        return np.stack(mus, -1), np.stack(vs, -1)

    def predict_ys(self, Xnew, num_inducing_samples: int = None, full_cov=False, full_output_cov=False):
        dists = self.predict_dists(Xnew, num_inducing_samples=num_inducing_samples, full_cov=full_cov,
                                   full_output_cov=full_output_cov)
        return dists.mean(), dists.variance()

    def likelihoods(self):
        return [expert.likelihood for expert in self.experts_list]

    def noise_variances(self):
        return [expert.likelihood.variance.numpy() for expert in self.experts_list]
