"""mogpe/training/metrics.py: NLPD / MAE / RMSE computed from model.predict_y (the conditionals run on the GPU; these
are O(N*K) reductions of the returned predictive distribution)."""
import numpy as np

from .._model_core import MixtureDistribution


def _log_prob(model, y_dist, Y):
    # legacy predict_y is a MixtureSameFamily over Normal -> per-output log_prob [N, F] (metrics.py:9-16 averages it);
    # the Keras flavour's tfd.Mixture of MVNDiag gives a joint [N]
    legacy = type(model).__module__.startswith("mogpe_b200.mixture_of_experts")
    return y_dist.log_prob(Y, joint_over_outputs=not legacy)


def negative_log_predictive_density(model, dataset):
    X, Y = dataset
    y_dist = model.predict_y(X)
    if isinstance(y_dist, MixtureDistribution):
        log_probs = _log_prob(model, y_dist, np.asarray(Y))
    else:  # (mean, scale) tuple, as the reference's fallback: tfd.Normal(y_dist[0], y_dist[1])
        z = (np.asarray(Y) - y_dist[0]) / y_dist[1]
        log_probs = -0.5 * z * z - np.log(np.abs(y_dist[1])) - 0.5 * np.log(2 * np.pi)
    return -float(np.mean(log_probs))


def mean_absolute_error(model, dataset):
    X, Y = dataset
    y_dist = model.predict_y(X)
    y_mean = y_dist.mean() if isinstance(y_dist, MixtureDistribution) else y_dist[0]
    return float(np.mean(np.abs(y_mean - np.asarray(Y))))


def root_mean_squared_error(model, dataset, batched=False):
    X, Y = dataset
    y_dist = model.predict_y(X)
    y_mean = y_dist.mean() if isinstance(y_dist, MixtureDistribution) else y_dist[0]
    return float(np.sqrt(np.mean((y_mean - np.asarray(Y)) ** 2)))


def _closure(fn, model, dataset):
    def closure():
        return fn(model=model, dataset=next(dataset))

    return closure


def build_negative_log_predictive_density(model, dataset, compile: bool = True):
    return _closure(negative_log_predictive_density, model, dataset)


def build_mean_absolute_error(model, dataset, compile: bool = True):
    return _closure(mean_absolute_error, model, dataset)


def build_root_mean_squared_error(model, dataset, compile: bool = True):
    return _closure(root_mean_squared_error, model, dataset)
