"""mogpe/gps/svgp.py: SVGPModel = gpflow SVGP + inducing-point sampling; legacy components subclass it."""
import numpy as np

from ..gpflow_lite import SVGP


class SVGPModel(SVGP):
    def __init__(self, kernel, likelihood, inducing_variable, mean_function=None, num_latent_gps: int = 1,
                 q_diag: bool = False, q_mu=None, q_sqrt=None, whiten: bool = True, num_data=None):
        super().__init__(kernel, likelihood, inducing_variable, mean_function=mean_function,
                         num_latent_gps=num_latent_gps, q_diag=q_diag, q_mu=q_mu, q_sqrt=q_sqrt, whiten=whiten,
                         num_data=num_data)
        self._owner = self
        self._parent_model = None

    @property
    def _model(self):
        return self._parent_model

    def _require_model(self):
        if self._parent_model is None:
            raise RuntimeError("attach the component to a MixtureOfSVGPExperts before calling predict methods "
                               "(the CUDA engine is built per model)")
        return self._parent_model

    def sample_inducing_points(self, num_samples: int = None, eps_u=None, rng=None) -> np.ndarray:
        """u ~ N(q_mu, q_sqrt q_sqrt^T) -> [S, M, L] (gps/svgp.py:54-74); a host-side convenience."""
        S = num_samples or 1
        M, L = self.q_mu.numpy().shape
        if eps_u is None:
            eps_u = (rng or np.random.default_rng()).standard_normal((S, L, M))
        u = self.q_mu.numpy().T[None] + np.einsum("lmk,slk->slm", self.q_sqrt_full(), np.asarray(eps_u))
        return np.transpose(u, (0, 2, 1))

    def predict_f(self, Xnew, num_inducing_samples: int = None, full_cov: bool = False, full_output_cov: bool = False,
                  eps_u=None):
        """(mean, var): [N, L], or [S, N, L] when sampling the inducing outputs (gps/svgp.py:76-152)."""
        if full_cov or full_output_cov:
            raise NotImplementedError("full_cov / full_output_cov are not on the accelerated path")
        m = self._require_model()
        if num_inducing_samples is None:
            return m._predict_f(self, Xnew)
        return m._predict_f_inducing(self, Xnew, num_inducing_samples, eps_u)

    def predict_y(self, Xnew, num_inducing_samples: int = None, full_cov: bool = False, full_output_cov: bool = False,
                  eps_u=None):
        f_mean, f_var = self.predict_f(Xnew, num_inducing_samples, full_cov, full_output_cov, eps_u)
        noise = np.broadcast_to(np.asarray(self.likelihood.variance.numpy()).reshape(-1), (self.num_latent_gps,))
        return f_mean, f_var + noise

    def prior_kl(self):
        return self._require_model()._prior_kl(self)
