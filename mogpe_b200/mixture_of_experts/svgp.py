"""mogpe/mixture_of_experts/svgp.py: legacy MixtureOfSVGPExperts on the CUDA engine."""
from typing import Optional, Tuple

import numpy as np

from .._model_core import Adam, MeanTracker, MixtureCore, MixtureDistribution
from ..experts import SVGPExperts
from ..gating_networks import SVGPGatingNetwork
from .base import MixtureOfExperts

BOUNDS = ("further_gating", "tight", "further")


class MixtureOfSVGPExperts(MixtureOfExperts):
    """(gating_network, experts, num_data=None, num_samples=1, bound="further_gating") as the reference
    (svgp.py:34-41).  Supported bounds: further_gating, tight, further; the experimental
    variants (tight_2, further_expert, ...) are out of scope and raise."""

    def __init__(self, gating_network: SVGPGatingNetwork, experts: SVGPExperts, num_data: Optional[int] = None,
                 num_samples: Optional[int] = 1, bound: Optional[str] = "further_gating", device: int = 0):
        assert isinstance(gating_network, SVGPGatingNetwork)
        assert isinstance(experts, SVGPExperts)
        super().__init__(gating_network, experts)
        self.num_samples, self.num_data, self.bound = num_samples, num_data, bound
        self.loss_tracker = MeanTracker(name="loss")
        self.optimizer = None
        for e in experts.experts_list:
            e._parent_model = self
        gating_network._parent_model = self
        self._core = MixtureCore(list(experts.experts_list), gating_network, flavour="legacy", device=device)

    @property
    def metrics(self):
        return [self.loss_tracker]

    @property
    def trainable_variables(self):
        return self._core.trainable_parameters

    def set_seed(self, seed: int):
        self._core.set_seed(seed)

    def set_fast_training(self, on: bool = True):
        """Gradient steps in the "1e-4 class" (six-slice INT8 products; see ModelCore.set_fast_training)."""
        self._core.set_fast_training(on)

    def set_precision(self, precision: str):
        """"float64" (default) or "tf32" (prediction on the tcgen05 tensor cores; new capability)."""
        self._core.set_precision(precision)

    def enable_data_parallel(self, rank: int, world_size: int, unique_id: bytes):
        self._core.enable_data_parallel(rank, world_size, unique_id)

    # -- objective ----------------------------------------------------------------------------------------------
    def _bound_name(self):
        if self.bound in BOUNDS:
            return self.bound
        if self.bound in ("tight_2", "further_expert"):
            raise NotImplementedError(f"bound '{self.bound}' is one of the reference's experimental variants "
                                      "(mixture_of_experts/svgp.py:471-1052) and is out of scope")
        # svgp.py:118-123: unknown values fall back to the tight bound (with a printed warning)
        print('Incorrect value passed so MixtureOfSVGPExperts.bound, should be "tight" or "further_gating" or '
              '"further". Using further_gating as default.')
        return "tight"

    def maximum_log_likelihood_objective(self, data: Tuple[np.ndarray, np.ndarray], eps=None) -> float:
        X, Y = data
        res, _ = self._core.elbo(X, Y, self._bound_name(), self.num_samples, self.num_data, eps=eps, want_grads=False)
        return res.elbo

    def _lb(self, data, bound, eps):
        X, Y = data
        return self._core.elbo(X, Y, bound, self.num_samples, self.num_data, eps=eps, want_grads=False)[0].elbo

    def lower_bound_further_gating(self, data, eps=None):
        return self._lb(data, "further_gating", eps)

    def lower_bound_tight(self, data, eps=None):
        return self._lb(data, "tight", eps)

    def lower_bound_further(self, data, eps=None):
        return self._lb(data, "further", eps)

    def elbo(self, data, eps=None):
        """svgp.py:1054-1061."""
        return self.maximum_log_likelihood_objective(data, eps)

    def training_loss(self, data, eps=None):
        """gpflow ExternalDataTrainingLossMixin: -(objective + log prior density); no priors in the reference."""
        return -self.maximum_log_likelihood_objective(data, eps)

    def training_loss_closure(self, data_or_iterator, compile=True):
        it = data_or_iterator
        if isinstance(it, tuple):
            return lambda: self.training_loss(it)
        return lambda: self.training_loss(next(it))

    def elbo_and_gradients(self, data, eps=None):
        X, Y = data
        res, grads = self._core.elbo(X, Y, self._bound_name(), self.num_samples, self.num_data, eps=eps, want_grads=True)
        return res.elbo, [grads.get(id(p)) for p in self.trainable_variables]

    def compile(self, optimizer=None, **_):
        self.optimizer = optimizer if optimizer is not None else Adam()

    def train_step(self, data, eps=None):
        """svgp.py:67-100."""
        if self.optimizer is None:
            self.compile()
        elbo, grads = self.elbo_and_gradients(data, eps)
        self.optimizer.apply_gradients([(None if g is None else -g, v) for g, v in zip(grads, self.trainable_variables)])
        self.loss_tracker.update_state(-elbo)
        return {"loss": self.loss_tracker.result()}

    def fit(self, X, Y, epochs=1, batch_size=32, shuffle=True, verbose=0, seed=0, device_data=False):
        """Keras-style loop (the legacy class is a tf.keras.Model too) with the optimiser step on the device."""
        if self.optimizer is None:
            self.compile()
        if not isinstance(self.optimizer, Adam):
            raise NotImplementedError("fit() runs the device-side Adam; call train_step in your own loop otherwise")
        return self._core.fit_device(X, Y, self._bound_name(), self.num_samples, self.num_data, self.optimizer, epochs,
                                     batch_size, shuffle, seed, self.loss_tracker, verbose, device_data)

    # -- prediction -----------------------------------------------------------------------------------------------
    def __call__(self, Xnew, training: Optional[bool] = False):
        return self.predict_y(Xnew)

    call = __call__

    def predict_y(self, Xnew, eps_mc=None, **kwargs) -> MixtureDistribution:
        """MixtureSameFamily(Categorical(probs), Normal(mus, sqrt(vars))) (base.py:90-115)."""
        if kwargs.get("num_inducing_samples") is not None:
            # base.py:90-115 with kwargs forwarded to the gating network and the experts: everything gets a leading sample
            # axis.  The conditionals given inducing samples run in the CUDA library (mogpe_predict_f_inducing_samples); the
            # mixture moments are an elementwise O(S N F K) combination of the returned arrays.
            S = int(kwargs["num_inducing_samples"])
            probs = self.gating_network.predict_mixing_probs(Xnew, num_inducing_samples=S, eps_mc=eps_mc, eps_u=kwargs.get("eps_u"),
                                                             rng=kwargs.get("rng"))                      # [S, N, K]
            mus, vs = self.experts.predict_ys(Xnew, num_inducing_samples=S)                               # [S, N, F, K] / [.., N, F, K]
            vs = np.broadcast_to(vs, mus.shape)
            w = probs[..., None, :]
            mean = (w * mus).sum(-1)
            var = (w * (vs + mus * mus)).sum(-1) - mean * mean
            return MixtureDistribution(probs, mus, np.sqrt(vs), mean, var)
        out = self._core.predict(Xnew, ("f_mean", "f_var", "mix_probs", "y_mean", "y_var"), eps_mc=eps_mc)
        sl = self._core.latent_slices()
        mus = np.stack([out["f_mean"][:, s] for s in sl[:-1]], -1)
        vs = np.stack([out["f_var"][:, s] for s in sl[:-1]], -1)
        nz = np.stack([np.broadcast_to(np.asarray(e.likelihood.variance.numpy()).reshape(-1), (e.num_latent_gps,))
                       for e in self.experts.experts_list], -1)
        return MixtureDistribution(out["mix_probs"], mus, np.sqrt(vs + nz[None]), out["y_mean"], out["y_var"])

    def predict_experts_fs(self, Xnew, **kwargs):
        return self.experts.predict_fs(Xnew, **kwargs)

    def predict_experts_ys(self, Xnew, **kwargs):
        return self.experts.predict_ys(Xnew, **kwargs)

    # -- component callbacks ------------------------------------------------------------------------------------------
    def _component_slice(self, comp):
        comps = list(self.experts.experts_list) + [self.gating_network]
        return self._core.latent_slices()[[id(c) for c in comps].index(id(comp))]

    def _predict_f(self, comp, Xnew):
        out = self._core.predict(Xnew, ("f_mean", "f_var"))
        s = self._component_slice(comp)
        return out["f_mean"][:, s], out["f_var"][:, s]

    def _predict_f_inducing(self, comp, Xnew, S, eps_u=None):
        N = np.shape(Xnew)[0]
        eng = self._core.ensure_engine(N)
        s = self._component_slice(comp)
        packed = None
        if eps_u is not None:
            packed = np.zeros((eng.P, S, eng.Mp))
            e = np.asarray(eps_u, dtype=np.float64)
            packed[s, :, :e.shape[2]] = np.transpose(e, (1, 0, 2))
        fm, fv = eng.predict_f_inducing_samples(Xnew, S, packed)
        return fm[:, :, s], np.broadcast_to(fv[None, :, s], fm[:, :, s].shape).copy()

    def _predict_mixing_probs(self, Xnew, eps_mc=None):
        return self._core.predict(Xnew, ("mix_probs",), eps_mc=eps_mc)["mix_probs"]

    def _prior_kl(self, comp):
        eng = self._core.ensure_engine(1)
        return float(eng.prior_kl()[self._component_slice(comp)].sum())
