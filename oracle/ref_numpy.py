"""CPU oracle (numpy, float64) for the mogpe hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference ships no numeric test for this path and its arithmetic lives in
un-vendored GPflow 2.3.1 (fork) / TensorFlow-Probability 0.14.1 / TensorFlow 2.6.3, none of which
is installable here.  This file restates the reference's call sequence and
the published GPflow/TFP formulas; it is pinned by analytic known answers,
by an independently written torch-autograd twin (oracle/ref_torch.py) and by finite differences
(tests/test_oracle_*.py), not by reference outputs.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (mogpe_b200/) never does.

A model is described by a plain ``spec`` dict (constrained-space values, all float64):

    spec = {
      "flavour": "keras" | "legacy",           # which API flavour's quirks apply
      "bound": "further_gating" | "tight" | "further",
      "num_samples": S, "num_data": None | int,
      "experts": [ gp, ... ],                   # K entries
      "gating": gp,                             # G = 1 (Bernoulli, K=2) or G = K (Softmax)
    }
    gp = {
      "Z": [M, D], "kernels": [ {"lengthscales": [D] or scalar, "variance": scalar[, "active_dims": column indices]}, ... ],
            # one kernel shared by all L latents, or L kernels (SeparateIndependent, shared Z)
      "q_mu": [M, L], "q_sqrt": [L, M, M] (lower-triangular) -- or [M, L] with "q_diag": True (gpflow's diagonal covariance) --,
      "mean_c": None | [L],
      "noise_variance": [L]   (experts only; Gaussian likelihood variance per output)
    }

Random draws are explicit inputs (RNG parity with TF's Philox stream is impossible):
    eps_u_experts: [K][S, L, M]   standard normals for u ~ q(u)  (mogpe/keras/gps.py:20-29)
    eps_u_gating:  [S, G, M]
    eps_h:         [S, N, G]      marginal gating samples (keras/gating_networks.py:188-199)
    eps_f:         [S, N, F, K]   marginal expert samples ("further" bound)
"""
import math

import numpy as np
from scipy.linalg import cholesky as _chol
from scipy.linalg import solve_triangular as _trsm
from scipy.special import erf

DEFAULT_JITTER = 1e-6  # gpflow.config.default_jitter()
INV_PROBIT_JITTER = 1e-3  # gpflow.likelihoods.utils.inv_probit
LOG_2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------------------------
# Appendix A.1: parameter transforms
# ----------------------------------------------------------------------------------------------
def softplus(x):
    return np.logaddexp(0.0, x)


def softplus_inverse(y):
    y = np.asarray(y, dtype=np.float64)
    return y + np.log(-np.expm1(-y))


def fill_triangular_index(n):
    """Index map of TFP FillTriangular(upper=False): x[idx[i, j]] lands at (i, j), i >= j.

    fill_triangular(x) = lower_part(reshape(concat(x[n:], reverse(x)), [n, n])).
    """
    m = n * (n + 1) // 2
    ar = np.arange(m)
    full = np.concatenate([ar[n:], ar[::-1]]).reshape(n, n)
    return full


def fill_triangular(x):
    """x: [..., n(n+1)/2] -> [..., n, n] lower-triangular (TFP ordering)."""
    x = np.asarray(x)
    m = x.shape[-1]
    n = int(round(math.sqrt(0.25 + 2.0 * m) - 0.5))
    idx = fill_triangular_index(n)
    return np.tril(x[..., idx])


def fill_triangular_inverse(L):
    L = np.asarray(L)
    n = L.shape[-1]
    idx = fill_triangular_index(n)
    out = np.zeros(L.shape[:-2] + (n * (n + 1) // 2,), dtype=L.dtype)
    ii, jj = np.tril_indices(n)
    out[..., idx[ii, jj]] = L[..., ii, jj]
    return out


# ----------------------------------------------------------------------------------------------
# Appendix A.2: SquaredExponential kernel in gpflow's expanded squared-distance form
# ----------------------------------------------------------------------------------------------
def _lengthscales(kern, D):
    ls = np.asarray(kern["lengthscales"], dtype=np.float64)
    return np.broadcast_to(ls, (D,)) if ls.ndim == 0 else ls


def _slice(kern, X):
    """gpflow Kernel.slice: a kernel built with active_dims sees only those input columns."""
    ad = kern.get("active_dims")
    return X if ad is None or X is None else X[..., np.asarray(ad, dtype=np.int64)]


def rbf_K(kern, X, X2=None):
    X, X2 = _slice(kern, X), _slice(kern, X2)
    D = X.shape[-1]
    ls = _lengthscales(kern, D)
    Xs = X / ls
    if X2 is None:
        sq = np.sum(Xs * Xs, -1, keepdims=True)
        dist = -2.0 * Xs @ Xs.T
        dist = dist + (sq + sq.T)
    else:
        X2s = X2 / ls
        sq1 = np.sum(Xs * Xs, -1)
        sq2 = np.sum(X2s * X2s, -1)
        dist = -2.0 * Xs @ X2s.T
        dist = dist + (sq1[:, None] + sq2[None, :])
    return float(kern["variance"]) * np.exp(-0.5 * dist)


def rbf_Kdiag(kern, X):
    return np.full((X.shape[0],), float(kern["variance"]))


def q_sqrt_full(gp):
    """[L, M, M] Cholesky factors of q(u); gpflow's q_diag=True stores their diagonals as [M, L] (SVGP._init_variational_parameters)."""
    q = np.asarray(gp["q_sqrt"], dtype=np.float64)
    if not gp.get("q_diag"):
        return q
    out = np.zeros((q.shape[1], q.shape[0], q.shape[0]))
    idx = np.arange(q.shape[0])
    out[:, idx, idx] = q.T
    return out


def _kernel_for_latent(gp, l):
    ks = gp["kernels"]
    return ks[0] if len(ks) == 1 else ks[l]


def num_latent(gp):
    return np.asarray(gp["q_mu"]).shape[1]


def mean_function(gp, N):
    L = num_latent(gp)
    c = gp.get("mean_c")
    if c is None:
        return np.zeros((N, L))
    return np.broadcast_to(np.asarray(c, dtype=np.float64).reshape(1, -1), (N, L)).copy()


# ----------------------------------------------------------------------------------------------
# Appendix A.3: whitened conditional, full_cov=False
# ----------------------------------------------------------------------------------------------
def base_conditional_latent(kern, X, Z, f, q_sqrt_l, jitter=DEFAULT_JITTER):
    """One latent: f [M, R] (R mean vectors), q_sqrt_l [M, M] or None -> mean [N, R], var [N]."""
    M = Z.shape[0]
    Kmm = rbf_K(kern, Z) + jitter * np.eye(M)
    Kmn = rbf_K(kern, Z, X)
    Lm = _chol(Kmm, lower=True)
    A = _trsm(Lm, Kmn, lower=True)
    fvar = rbf_Kdiag(kern, X) - np.sum(A * A, 0)
    fmean = A.T @ f
    if q_sqrt_l is not None:
        LTA = np.tril(q_sqrt_l).T @ A
        fvar = fvar + np.sum(LTA * LTA, 0)
    return fmean, fvar


def predict_f(gp, X):
    """gpflow SVGP.predict_f(full_cov=False): marginal q(f) -> mean [N, L], var [N, L]."""
    N = X.shape[0]
    L = num_latent(gp)
    Z = np.asarray(gp["Z"])
    mean = np.zeros((N, L))
    var = np.zeros((N, L))
    for l in range(L):
        m, v = base_conditional_latent(
            _kernel_for_latent(gp, l), X, Z, np.asarray(gp["q_mu"])[:, l : l + 1], q_sqrt_full(gp)[l]
        )
        mean[:, l] = m[:, 0]
        var[:, l] = v
    return mean + mean_function(gp, N), var


def sample_inducing(gp, eps_u):
    """u_s = q_mu[:, l] + tril(q_sqrt_l) @ eps[s, l]  (TFP MultivariateNormalTriL.sample). -> [S, M, L]."""
    q_mu = np.asarray(gp["q_mu"])
    q_sqrt = q_sqrt_full(gp)
    S, L, M = eps_u.shape
    u = np.zeros((S, M, L))
    for s in range(S):
        for l in range(L):
            u[s, :, l] = q_mu[:, l] + np.tril(q_sqrt[l]) @ eps_u[s, l]
    return u


def predict_f_given_inducing_samples(gp, X, eps_u):
    """mogpe/keras/gps.py:14-50 and mogpe/gps/svgp.py:130-152: per sample, conditional with q_sqrt=None.

    -> mean [S, N, L], var [S, N, L] (var identical for every s: no q_sqrt term, quirk 5)."""
    N = X.shape[0]
    L = num_latent(gp)
    Z = np.asarray(gp["Z"])
    u = sample_inducing(gp, eps_u)
    S = u.shape[0]
    mean = np.zeros((S, N, L))
    var = np.zeros((S, N, L))
    for s in range(S):
        for l in range(L):
            m, v = base_conditional_latent(_kernel_for_latent(gp, l), X, Z, u[s, :, l : l + 1], None)
            mean[s, :, l] = m[:, 0]
            var[s, :, l] = v
    return mean + mean_function(gp, N)[None], var


def prior_kl(gp):
    """Appendix A.5 (gauss_kl, whitened, full q_sqrt), summed over the L latents."""
    q_mu = np.asarray(gp["q_mu"])
    q_sqrt = np.tril(q_sqrt_full(gp))
    M, L = q_mu.shape
    mahal = np.sum(q_mu * q_mu)
    logdet = np.sum(np.log(np.square(np.diagonal(q_sqrt, axis1=-2, axis2=-1))))
    trace = np.sum(q_sqrt * q_sqrt)
    return 0.5 * (mahal - M * L - logdet + trace)


# ----------------------------------------------------------------------------------------------
# Appendix A.6 / A.7: likelihood pieces and densities
# ----------------------------------------------------------------------------------------------
def inv_probit(x):
    return 0.5 * (1.0 + erf(x / math.sqrt(2.0))) * (1.0 - 2.0 * INV_PROBIT_JITTER) + INV_PROBIT_JITTER


def softmax(x):
    x = x - np.max(x, -1, keepdims=True)
    e = np.exp(x)
    return e / np.sum(e, -1, keepdims=True)


def normal_prob(y, loc, scale):
    z = (y - loc) / scale
    return np.exp(-0.5 * z * z) / (scale * math.sqrt(2.0 * math.pi))


def normal_log_prob(y, loc, scale):
    z = (y - loc) / scale
    return -0.5 * z * z - np.log(scale) - 0.5 * LOG_2PI


def _noise(gp):
    L = num_latent(gp)
    return np.broadcast_to(np.asarray(gp["noise_variance"], dtype=np.float64).reshape(-1), (L,))


def _scale(num_data, N):
    return 1.0 if num_data is None else num_data / N


# ----------------------------------------------------------------------------------------------
# Gating network
# ----------------------------------------------------------------------------------------------
def _mixing_probs_given_h(spec, h):
    """conditional_mean path. h: [..., G]. Keras: keras/gating_networks.py:197-198,245-246;
    legacy: gating_networks/svgp.py:100-105."""
    G = h.shape[-1]
    if G == 1:
        if spec["flavour"] == "keras":
            return inv_probit(np.concatenate([h, -h], -1))
        p = inv_probit(h)
        return np.concatenate([p, 1.0 - p], -1)
    return softmax(h)


def _mixing_probs_given_h_meanvar(spec, h_mean, h_var, eps_mc=None):
    """predict_mean_and_var path (Appendix A.6). Bernoulli analytic; Softmax = MC over eps_mc [R, N, K]."""
    G = h_mean.shape[-1]
    if G == 1:
        if spec["flavour"] == "keras":
            hm = np.concatenate([h_mean, -h_mean], -1)
            hv = np.concatenate([h_var, h_var], -1)
            return inv_probit(hm / np.sqrt(1.0 + hv))
        p = inv_probit(h_mean / np.sqrt(1.0 + h_var))
        return np.concatenate([p, 1.0 - p], -1)
    if eps_mc is None:
        raise ValueError("Softmax predict_mean_and_var is Monte-Carlo: pass eps_mc [R, N, K]")
    samples = h_mean[None] + np.sqrt(h_var)[None] * eps_mc
    return np.mean(softmax(samples), 0)


def predict_mixing_probs(spec, X, eps_mc=None):
    """keras/gating_networks.py:173-186 ; legacy gating_networks/svgp.py:69-83 -> [N, K]."""
    h_mean, h_var = predict_f(spec["gating"], X)
    return _mixing_probs_given_h_meanvar(spec, h_mean, h_var, eps_mc)


# ----------------------------------------------------------------------------------------------
# ELBOs
# ----------------------------------------------------------------------------------------------
def _expert_probs_inducing(spec, X, Y, eps_u_experts):
    """p_k[s, n] = prod_f N(Y[n, f]; y_mean[s, n, f], scale)  (a5, a6)."""
    K = len(spec["experts"])
    S = eps_u_experts[0].shape[0]
    N = X.shape[0]
    out = np.zeros((S, N, K))
    for k, gp in enumerate(spec["experts"]):
        f_mean, f_var = predict_f_given_inducing_samples(gp, X, eps_u_experts[k])
        y_var = f_var + _noise(gp)[None, None, :]
        scale = y_var if spec["flavour"] == "keras" else np.sqrt(y_var)  # quirk 1
        out[:, :, k] = np.prod(normal_prob(Y[None], f_mean, scale), -1)
    return out


def _marginalise(experts_probs, mixing_probs, num_data):
    """a9: log sum_k p_k[s1, n] pi_k[s2, n] in probability space, mean over S x S, sum over n."""
    N = experts_probs.shape[1]
    marg = np.einsum("ank,bnk->abn", experts_probs, mixing_probs)
    var_exp = np.sum(np.mean(np.log(marg), (0, 1)))
    return var_exp * _scale(num_data, N)


def _kls(spec):
    return prior_kl(spec["gating"]), sum(prior_kl(gp) for gp in spec["experts"])


def _gating_h_samples(spec, X, eps_h):
    h_mean, h_var = predict_f(spec["gating"], X)
    sd = np.sqrt(h_var) if spec["flavour"] == "keras" else h_var  # quirk 2
    return h_mean[None] + sd[None] * eps_h


def lower_bound_further_gating(spec, X, Y, eps_u_experts, eps_h):
    """keras mosvgpe.py:109-168 ; legacy mixture_of_experts/svgp.py:129-230."""
    kl_g, kl_e = _kls(spec)
    mixing = _mixing_probs_given_h(spec, _gating_h_samples(spec, X, eps_h))
    experts = _expert_probs_inducing(spec, X, Y, eps_u_experts)
    return _marginalise(experts, mixing, spec.get("num_data")) - kl_g - kl_e


def lower_bound_tight(spec, X, Y, eps_u_experts, eps_u_gating, eps_mc=None):
    """keras mosvgpe.py:170-221 ; legacy svgp.py:332-469."""
    kl_g, kl_e = _kls(spec)
    h_mean, h_var = predict_f_given_inducing_samples(spec["gating"], X, eps_u_gating)
    S = h_mean.shape[0]
    mixing = np.stack(
        [_mixing_probs_given_h_meanvar(spec, h_mean[s], h_var[s], None if eps_mc is None else eps_mc[s]) for s in range(S)]
    )
    experts = _expert_probs_inducing(spec, X, Y, eps_u_experts)
    return _marginalise(experts, mixing, spec.get("num_data")) - kl_g - kl_e


def lower_bound_further(spec, X, Y, eps_f, eps_h):
    """keras mosvgpe.py:223-274 ; legacy svgp.py:232-330."""
    kl_g, kl_e = _kls(spec)
    mixing = _mixing_probs_given_h(spec, _gating_h_samples(spec, X, eps_h))  # [S, N, K]
    K = len(spec["experts"])
    S, N = eps_h.shape[0], X.shape[0]
    F = Y.shape[1]
    if spec["flavour"] == "keras":
        experts = np.zeros((S, N, K))
        for k, gp in enumerate(spec["experts"]):
            f_mean, f_var = predict_f(gp, X)
            f_s = f_mean[None] + np.sqrt(f_var)[None] * eps_f[..., k]
            noise = _noise(gp)[None, None, :]  # conditional_variance -> used as scale (quirk 2)
            experts[:, :, k] = np.prod(normal_prob(Y[None], f_s, noise), -1)
        return _marginalise(experts, mixing, spec.get("num_data")) - kl_g - kl_e
    # legacy: per-output mixture log_prob, same sample index for gating and experts
    logp = np.zeros((S, N, F, K))
    for k, gp in enumerate(spec["experts"]):
        f_mean, f_var = predict_f(gp, X)
        f_s = f_mean[None] + f_var[None] * eps_f[..., k]  # variance used as scale
        logp[..., k] = normal_log_prob(Y[None], f_s, _noise(gp)[None, None, :])
    w = logp + np.log(mixing)[:, :, None, :]
    mx = np.max(w, -1, keepdims=True)
    lse = mx[..., 0] + np.log(np.sum(np.exp(w - mx), -1))  # tfd.Mixture.log_prob
    var_exp = np.sum(np.mean(np.sum(lse, -1), 0))
    return var_exp * _scale(spec.get("num_data"), N) - kl_g - kl_e


def elbo(spec, X, Y, eps):
    b = spec["bound"]
    if b == "further_gating":
        return lower_bound_further_gating(spec, X, Y, eps["u_experts"], eps["h"])
    if b == "tight":
        return lower_bound_tight(spec, X, Y, eps["u_experts"], eps["u_gating"], eps.get("mc"))
    if b == "further":
        return lower_bound_further(spec, X, Y, eps["f"], eps["h"])
    raise ValueError(b)


# ----------------------------------------------------------------------------------------------
# Prediction (a11, a12)
# ----------------------------------------------------------------------------------------------
def predict_experts_y(spec, X):
    """-> y_mean [N, F, K], y_var [N, F, K]."""
    mus, vs = [], []
    for gp in spec["experts"]:
        f_mean, f_var = predict_f(gp, X)
        mus.append(f_mean)
        vs.append(f_var + _noise(gp)[None, :])
    return np.stack(mus, -1), np.stack(vs, -1)


def predict_y(spec, X, eps_mc=None):
    """Mixture mean / variance [N, F] (Appendix A.7).

    keras: tfd.Mixture of MVNDiag(loc, scale_diag=y_var) (keras/mixture_of_experts/base.py:65-77,
    experts.py:167); legacy: MixtureSameFamily(Normal(mu, sqrt(var))) (mixture_of_experts/base.py:90-115)."""
    pi = predict_mixing_probs(spec, X, eps_mc)  # [N, K]
    mu, var = predict_experts_y(spec, X)  # [N, F, K]
    w = pi[:, None, :]
    mean = np.sum(w * mu, -1)
    if spec["flavour"] == "keras":
        sd = var  # quirk 1: component stddev is y_var
        mix_var = np.sum(w * sd * sd, -1) + np.sum(w * mu * mu, -1) - mean * mean
        mix_var = np.square(np.sqrt(mix_var))  # Distribution.variance() falls back to stddev()**2
    else:
        mix_var = np.sum(w * var, -1) + np.sum(w * np.square(mu - mean[..., None]), -1)
    return mean, mix_var
