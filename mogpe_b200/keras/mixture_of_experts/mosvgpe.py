"""mogpe/keras/mixture_of_experts/mosvgpe.py: MixtureOfSVGPExperts (Keras flavour) on the CUDA engine."""
from typing import List, Optional

import numpy as np

from ..._model_core import Adam, MeanTracker, MixtureCore, MixtureDistribution
from ..experts import SVGPExpert
from ..gating_networks import SVGPGatingNetwork
from .base import MixtureOfExpertsBase

BOUNDS = ("further_gating", "tight", "further")


def variational_expectation_scale(num_data, batch_size):
    """mosvgpe.py:309-316."""
    return num_data / batch_size if num_data is not None else 1.0


class MixtureOfSVGPExperts(MixtureOfExpertsBase):
    """Mixture of SVGP experts trained with stochastic variational inference (mosvgpe.py:18-306).

    Same constructor, bounds, train_step / test_step, predict_y and call as the reference.  Differences forced
    by the environment: Monte-Carlo draws come from the library's Philox stream (or explicit `eps`), not TF's;
    returned distributions are light numpy objects, not tfd instances."""

    def __init__(self, experts_list: List[SVGPExpert], gating_network: SVGPGatingNetwork, num_samples: Optional[int] = 1,
                 num_data: Optional[int] = None, bound: Optional[str] = "further_gating", name: str = "MoSVGPE",
                 device: int = 0):
        for expert in experts_list:
            assert isinstance(expert, SVGPExpert)
        assert isinstance(gating_network, SVGPGatingNetwork)
        super().__init__(experts_list=experts_list, gating_network=gating_network, name=name)
        self._num_data, self._num_samples, self._bound = num_data, num_samples, bound
        self.loss_tracker = MeanTracker(name="loss")
        self.optimizer = None
        for e in experts_list:
            e._parent_model = self
        gating_network._parent_model = self
        self._core = MixtureCore([e.gp for e in experts_list], gating_network.gp, flavour="keras", device=device)

    # -- reference properties ------------------------------------------------------------------------------
    @property
    def metrics(self):
        return [self.loss_tracker]

    @property
    def num_data(self):
        return self._num_data

    @property
    def num_samples(self):
        return self._num_samples

    @property
    def bound(self):
        return self._bound

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
        """Data-parallel fit(): see MixtureCore.enable_data_parallel (new capability; the reference is single-device)."""
        self._core.enable_data_parallel(rank, world_size, unique_id)

    # -- objective ---------------------------------------------------------------------------------------------
    def maximum_log_likelihood_objective(self, data, eps=None) -> float:
        return self.elbo(data=data, num_samples=self.num_samples, num_data=self.num_data, bound=self.bound, eps=eps)

    def elbo(self, data, num_samples: Optional[int] = 1, num_data: Optional[int] = None,
             bound: Optional[str] = "further_gating", eps=None) -> float:
        """mosvgpe.py:83-107 (unknown bound -> further_gating, with the reference's printed warning)."""
        if bound not in BOUNDS:
            print('bound should be "tight" or "further_gating" or "further". Using further_gating as default.')
            bound = "further_gating"
        X, Y = data
        res, _ = self._core.elbo(X, Y, bound, num_samples, num_data, eps=eps, want_grads=False)
        return res.elbo

    def lower_bound_further_gating(self, data_batch, num_data: Optional[int] = 1, num_samples: Optional[int] = 1, eps=None):
        return self.elbo(data_batch, num_samples, num_data, "further_gating", eps)

    def lower_bound_tight(self, data_batch, num_data: Optional[int] = 1, num_samples: Optional[int] = 1, eps=None):
        return self.elbo(data_batch, num_samples, num_data, "tight", eps)

    def lower_bound_further(self, data_batch, num_data: Optional[int] = 1, num_samples: Optional[int] = 1, eps=None):
        return self.elbo(data_batch, num_samples, num_data, "further", eps)

    def elbo_and_gradients(self, data, eps=None):
        """-> (elbo, [d elbo / d v for v in trainable_variables]) w.r.t. the unconstrained variables: what
        tape.gradient(-loss, trainable_variables) gives the reference (mosvgpe.py:58-62), negated."""
        X, Y = data
        res, grads = self._core.elbo(X, Y, self.bound if self.bound in BOUNDS else "further_gating",
                                     self.num_samples, self.num_data, eps=eps, want_grads=True)
        return res.elbo, [grads.get(id(p)) for p in self.trainable_variables]

    # -- training (Keras loop) -----------------------------------------------------------------------------
    def compile(self, optimizer=None, **_):
        self.optimizer = optimizer if optimizer is not None else Adam()

    def train_step(self, data, eps=None):
        """mosvgpe.py:57-69: loss = -ELBO, gradients w.r.t. trainable_variables, optimizer.apply_gradients."""
        if self.optimizer is None:
            self.compile()
        elbo, grads = self.elbo_and_gradients(data, eps)
        loss = -elbo
        self.optimizer.apply_gradients([(None if g is None else -g, v) for g, v in zip(grads, self.trainable_variables)])
        self.loss_tracker.update_state(loss)
        return {"loss": self.loss_tracker.result()}

    def test_step(self, data, eps=None):
        return {"loss": -self.maximum_log_likelihood_objective(data, eps)}

    def fit(self, X, Y, epochs=1, batch_size=32, shuffle=True, verbose=0, seed=0, device_optimizer=True,
            device_data=False):
        """Minimal tf.keras.Model.fit: minibatches (last partial batch kept, as Keras does), per-epoch mean loss.

        device_optimizer=True (default, Adam only) keeps the variables, the bijector chain and Adam on the GPU for the
        whole loop and copies the variables back into the Parameter objects at the end; False runs train_step per
        batch (gradients and Adam on the host, as the reference's eager loop).  device_data=True additionally keeps the
        data set in HBM and assembles every shuffled minibatch with a device gather (no per-step H2D)."""
        if self.optimizer is None:
            self.compile()
        if device_optimizer and isinstance(self.optimizer, Adam):
            bound = self.bound if self.bound in BOUNDS else "further_gating"
            return self._core.fit_device(X, Y, bound, self.num_samples, self.num_data, self.optimizer, epochs, batch_size,
                                         shuffle, seed, self.loss_tracker, verbose, device_data)
        X, Y = np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64)
        rng = np.random.default_rng(seed)
        history = {"loss": []}
        n = X.shape[0]
        for epoch in range(epochs):
            self.loss_tracker.reset_states()
            order = rng.permutation(n) if shuffle else np.arange(n)
            for s in range(0, n, batch_size):
                idx = order[s:s + batch_size]
                logs = self.train_step((np.ascontiguousarray(X[idx]), np.ascontiguousarray(Y[idx])))
            history["loss"].append(logs["loss"])
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs} - loss: {logs['loss']:.4f}")
        return history

    # -- prediction ----------------------------------------------------------------------------------------------
    def __call__(self, Xnew, training: Optional[bool] = False):
        return self.call(Xnew, training)

    def call(self, Xnew, training: Optional[bool] = False):
        if not training:
            dist = self.predict_y(Xnew)
            return dist.mean(), dist.variance()

    def predict_y(self, Xnew, eps_mc=None, **kwargs) -> MixtureDistribution:
        """tfd.Mixture(Categorical(mixing probs), [MVNDiag(y_mean_k, scale_diag=y_var_k)]) (base.py:65-77)."""
        out = self._core.predict(Xnew, ("f_mean", "f_var", "mix_probs", "y_mean", "y_var"), eps_mc=eps_mc)
        mus, vs = self._stack_experts(out, noise=True)
        return MixtureDistribution(out["mix_probs"], mus, vs, out["y_mean"], out["y_var"])  # scale = variance (quirk 1)

    def _stack_experts(self, out, noise):
        sl = self._core.latent_slices()
        mus = np.stack([out["f_mean"][:, s] for s in sl[:-1]], -1)
        vs = np.stack([out["f_var"][:, s] for s in sl[:-1]], -1)
        if noise:
            nz = np.stack([np.broadcast_to(np.asarray(e.gp.likelihood.variance.numpy()).reshape(-1), (e.gp.num_latent_gps,))
                           for e in self.experts_list], -1)
            vs = vs + nz[None]
        return mus, vs

    def predict_experts_f(self, Xnew):
        """-> f_means, f_vars stacked on the last axis [N, F, K] (mosvgpe.py:276-282)."""
        return self._stack_experts(self._core.predict(Xnew, ("f_mean", "f_var")), noise=False)

    def predict_experts_y(self, Xnew):
        return self._stack_experts(self._core.predict(Xnew, ("f_mean", "f_var")), noise=True)

    # -- component callbacks ---------------------------------------------------------------------------------------
    def _component_slice(self, comp):
        comps = list(self.experts_list) + [self.gating_network]
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
            e = np.asarray(eps_u, dtype=np.float64)  # [S, L, M]
            packed[s, :, :e.shape[2]] = np.transpose(e, (1, 0, 2))
        fm, fv = eng.predict_f_inducing_samples(Xnew, S, packed)
        return fm[:, :, s], np.broadcast_to(fv[None, :, s], fm[:, :, s].shape).copy()

    def _predict_mixing_probs(self, Xnew, eps_mc=None):
        return self._core.predict(Xnew, ("mix_probs",), eps_mc=eps_mc)["mix_probs"]

    def _prior_kl(self, comp):
        eng = self._core.ensure_engine(1)
        return float(eng.prior_kl()[self._component_slice(comp)].sum())
