"""Minimal stand-ins for the GPflow 2.3 objects that the reference's constructors take (kernels, likelihoods,
mean functions, inducing variables, Parameter + bijectors).  They hold parameters only: every number the
hot path computes from them is computed by the CUDA library.

Mirrors (reference call sites): gpflow.kernels.RBF / SquaredExponential, SeparateIndependent, SharedIndependent
(mogpe/keras/gating_networks.py:99-106), gpflow.likelihoods.Gaussian / Bernoulli / Softmax
(mogpe/gps/likelihoods.py:12-67), gpflow.mean_functions.Constant / Zero, gpflow.inducing_variables.InducingPoints,
gpflow.base.Parameter with gpflow.utilities.positive / triangular.
"""
import math
from typing import List, Optional, Sequence

import numpy as np

DEFAULT_FLOAT = np.float64
DEFAULT_JITTER = 1e-6


# ---- bijectors -------------------------------------------------------------------------------------------
class Identity:
    def forward(self, u):
        return u

    def inverse(self, x):
        return x

    def chain_grad(self, u, g_constrained):
        return g_constrained


class Positive:
    """gpflow.utilities.positive(lower): x = lower + softplus(u)."""

    def __init__(self, lower=0.0):
        self.lower = float(lower)

    def forward(self, u):
        return self.lower + np.logaddexp(0.0, u)

    def inverse(self, x):
        y = np.asarray(x, dtype=DEFAULT_FLOAT) - self.lower
        if np.any(y <= 0):
            raise ValueError(f"value must be > {self.lower} for a positive parameter")
        return y + np.log(-np.expm1(-y))

    def chain_grad(self, u, g_constrained):
        return g_constrained / (1.0 + np.exp(-u))  # d softplus / du = sigmoid(u)


def fill_triangular_index(n):
    """TFP FillTriangular(upper=False): x[idx[i, j]] lands at (i, j) for i >= j."""
    m = n * (n + 1) // 2
    ar = np.arange(m)
    return np.concatenate([ar[n:], ar[::-1]]).reshape(n, n)


class Triangular:
    """gpflow.utilities.triangular() = tfp.bijectors.FillTriangular: [..., n(n+1)/2] <-> lower-triangular [..., n, n]."""

    def forward(self, u):
        m = u.shape[-1]
        n = int(round(math.sqrt(0.25 + 2.0 * m) - 0.5))
        return np.tril(u[..., fill_triangular_index(n)])

    def inverse(self, x):
        x = np.asarray(x, dtype=DEFAULT_FLOAT)
        n = x.shape[-1]
        idx = fill_triangular_index(n)
        ii, jj = np.tril_indices(n)
        out = np.zeros(x.shape[:-2] + (n * (n + 1) // 2,), dtype=DEFAULT_FLOAT)
        out[..., idx[ii, jj]] = x[..., ii, jj]
        return out

    def chain_grad(self, u, g_constrained):
        n = g_constrained.shape[-1]
        idx = fill_triangular_index(n)
        ii, jj = np.tril_indices(n)
        out = np.zeros_like(u)
        out[..., idx[ii, jj]] = g_constrained[..., ii, jj]
        return out


def positive(lower=0.0):
    return Positive(lower)


def triangular():
    return Triangular()


class Parameter:
    """gpflow.Parameter: stores the UNCONSTRAINED value (what the optimiser updates, as the reference's
    trainable_variables are) and exposes the constrained one."""

    _clock = [0]  # global modification counter: lets the engine skip re-uploading parameters that did not change
    # Device-side training (fit) leaves the freshest variables in HBM and registers a callback here instead of copying them
    # back right away; the first READ of any Parameter value runs the callbacks (one D2H copy), so a sequence of fit() calls
    # never pays for it and every other reader still sees the trained values.
    _pending_pulls = []

    @classmethod
    def flush_pending(cls):
        while cls._pending_pulls:
            cls._pending_pulls.pop()()

    def __init__(self, value, transform=None, trainable=True, name=None):
        self.transform = transform or Identity()
        self.unconstrained = np.array(self.transform.inverse(np.asarray(value, dtype=DEFAULT_FLOAT)), dtype=DEFAULT_FLOAT)
        self.trainable = trainable
        self.name = name

    @property
    def unconstrained(self):
        if Parameter._pending_pulls:
            Parameter.flush_pending()
        return self._u

    @unconstrained.setter
    def unconstrained(self, u):
        if Parameter._pending_pulls:
            Parameter.flush_pending()  # a deferred copy-back must not overwrite this assignment later
        u = np.array(u, dtype=DEFAULT_FLOAT)
        u.flags.writeable = False  # every change goes through this setter (and bumps `version`)
        self._u = u
        Parameter._clock[0] += 1
        self.version = Parameter._clock[0]

    def numpy(self):
        return self.transform.forward(self.unconstrained)

    value = numpy

    @property
    def shape(self):
        return self.numpy().shape

    def assign(self, value):
        self.unconstrained = np.array(self.transform.inverse(np.asarray(value, dtype=DEFAULT_FLOAT)), dtype=DEFAULT_FLOAT)

    def assign_unconstrained(self, u):
        self.unconstrained = np.array(u, dtype=DEFAULT_FLOAT).reshape(self.unconstrained.shape)

    def chain_grad(self, g_constrained):
        g = np.asarray(g_constrained, dtype=DEFAULT_FLOAT).reshape(self.numpy().shape)
        return self.transform.chain_grad(self.unconstrained, g)

    def __repr__(self):
        return f"Parameter({self.name}, shape={self.unconstrained.shape}, trainable={self.trainable})"


class Module:
    """Collects Parameters through attributes, lists and sub-modules (like tf.Module.trainable_variables)."""

    def _children(self):
        for k, v in vars(self).items():
            if k.startswith("_parent") or k.startswith("_owner"):
                continue  # back-references (component -> model), not children
            yield k, v

    def parameters(self, prefix=""):
        out = []
        for k, v in self._children():
            name = f"{prefix}{k.lstrip('_')}"
            if isinstance(v, Parameter):
                out.append((name, v))
            elif isinstance(v, Module):
                out += v.parameters(name + "/")
            elif isinstance(v, (list, tuple)):
                for i, e in enumerate(v):
                    if isinstance(e, Parameter):
                        out.append((f"{name}/{i}", e))
                    elif isinstance(e, Module):
                        out += e.parameters(f"{name}/{i}/")
        return out

    @property
    def trainable_parameters(self):
        seen, out = set(), []
        for _, p in self.parameters():
            if p.trainable and id(p) not in seen:
                seen.add(id(p))
                out.append(p)
        return out


# ---- kernels ---------------------------------------------------------------------------------------------
class Kernel(Module):
    pass


class SquaredExponential(Kernel):
    """k(x, z) = variance * exp(-0.5 |x/l - z/l|^2); lengthscales scalar or [D] (ARD)."""

    def __init__(self, variance=1.0, lengthscales=1.0, active_dims=None, name=None):
        self.variance = Parameter(variance, transform=positive(), name="variance")
        self.lengthscales = Parameter(lengthscales, transform=positive(), name="lengthscales")
        self.name = name
        # gpflow.kernels.Kernel(active_dims=...): the kernel acts on these input columns only (a slice or a list of column
        # indices; None = all).  The CUDA path keeps the full-width layout and gives the other columns an effectively infinite
        # lengthscale (engine.INACTIVE_LENGTHSCALE), which is the same function and has the same gradients.
        self.active_dims = active_dims
        if active_dims is not None and not isinstance(active_dims, slice):
            ad = np.asarray(active_dims, dtype=np.int64).reshape(-1)
            if ad.size == 0 or np.any(ad < 0) or len(set(ad.tolist())) != ad.size:
                raise ValueError("active_dims must be distinct non-negative column indices")
            if self.lengthscales.unconstrained.ndim > 0 and self.lengthscales.unconstrained.size != ad.size:
                raise ValueError("ARD lengthscales need one entry per active dimension")
            self.active_dims = ad

    def active_columns(self, input_dim: int):
        """Column indices this kernel reads for inputs with `input_dim` columns (None: all of them)."""
        ad = self.active_dims
        if ad is None:
            return None
        cols = np.arange(input_dim)[ad] if isinstance(ad, slice) else np.asarray(ad)
        if cols.size == 0 or cols.max() >= input_dim:
            raise ValueError(f"active_dims {ad} do not fit inputs with {input_dim} columns")
        return None if (cols.size == input_dim and np.array_equal(cols, np.arange(input_dim))) else cols

    @property
    def ard(self):
        return self.lengthscales.unconstrained.ndim > 0


RBF = SquaredExponential


class MultioutputKernel(Kernel):
    pass


class SeparateIndependent(MultioutputKernel):
    def __init__(self, kernels: Sequence[Kernel], name=None):
        self.kernels = list(kernels)
        for k in self.kernels:
            if not isinstance(k, SquaredExponential):
                raise NotImplementedError("only SquaredExponential kernels are supported by the CUDA path")
        self.name = name

    @property
    def num_latent_gps(self):
        return len(self.kernels)

    @property
    def latent_kernels(self):
        return tuple(self.kernels)


class SharedIndependent(MultioutputKernel):
    def __init__(self, kernel: Kernel, output_dim: int, name=None):
        if not isinstance(kernel, SquaredExponential):
            raise NotImplementedError("only SquaredExponential kernels are supported by the CUDA path")
        self.kernel = kernel
        self.output_dim = int(output_dim)
        self.name = name

    @property
    def num_latent_gps(self):
        return self.output_dim

    @property
    def latent_kernels(self):
        return (self.kernel,)


def latent_kernels(kernel, num_latent_gps) -> List[SquaredExponential]:
    """-> one shared kernel, or one kernel per latent GP."""
    if isinstance(kernel, SeparateIndependent):
        if kernel.num_latent_gps != num_latent_gps:
            raise ValueError("SeparateIndependent needs one kernel per latent GP")
        return list(kernel.kernels)
    if isinstance(kernel, SharedIndependent):
        return [kernel.kernel]
    if isinstance(kernel, SquaredExponential):
        return [kernel]
    raise NotImplementedError(f"kernel {type(kernel).__name__} is not supported by the CUDA path "
                              "(SquaredExponential / RBF, SeparateIndependent, SharedIndependent)")


# ---- likelihoods -----------------------------------------------------------------------------------------
class Likelihood(Module):
    pass


class Gaussian(Likelihood):
    """mogpe/gps/likelihoods.py:12-67 (and gpflow's): variance > variance_lower_bound, scalar or [1, F]."""

    DEFAULT_VARIANCE_LOWER_BOUND = 1e-6

    def __init__(self, variance=1.0, variance_lower_bound=DEFAULT_VARIANCE_LOWER_BOUND, name=None):
        self.variance_lower_bound = variance_lower_bound
        self.variance = Parameter(variance, transform=positive(lower=variance_lower_bound), name="variance")
        self.name = name


def inv_probit(x):
    """gpflow.likelihoods.utils.inv_probit: 0.5 (1 + erf(x / sqrt 2)) (1 - 2e-3) + 1e-3."""
    from math import sqrt

    try:
        from scipy.special import erf
    except ImportError:  # pragma: no cover
        erf = np.vectorize(math.erf)
    return 0.5 * (1.0 + erf(np.asarray(x, dtype=DEFAULT_FLOAT) / sqrt(2.0))) * (1.0 - 2e-3) + 1e-3


class Bernoulli(Likelihood):
    """Probit link with gpflow's 1e-3 jitter.  The two maps below act elementwise on gating-function
    values the caller already holds on the host (predict_mixing_probs_given_h); everything that starts from inputs X runs
    in the CUDA library."""

    def __init__(self, name=None):
        self.name = name

    def conditional_mean(self, F):
        return inv_probit(F)

    def predict_mean_and_var(self, Fmu, Fvar):
        """gpflow Bernoulli.predict_mean_and_var (probit link, analytic): p = inv_probit(mu / sqrt(1 + var)), p - p^2."""
        p = inv_probit(np.asarray(Fmu, dtype=DEFAULT_FLOAT) / np.sqrt(1.0 + np.asarray(Fvar, dtype=DEFAULT_FLOAT)))
        return p, p - p * p


class Softmax(Likelihood):
    def __init__(self, num_classes, name=None):
        self.num_classes = int(num_classes)
        self.num_monte_carlo_points = 100  # gpflow MonteCarloLikelihood default
        self.name = name

    def conditional_mean(self, F):
        F = np.asarray(F, dtype=DEFAULT_FLOAT)
        e = np.exp(F - F.max(-1, keepdims=True))
        return e / e.sum(-1, keepdims=True)

    def predict_mean_and_var(self, Fmu, Fvar, eps=None, rng=None):
        """gpflow MonteCarloLikelihood.predict_mean_and_var: mean (and variance) of softmax(mu + sqrt(var) eps) over
        num_monte_carlo_points standard-normal draws; eps [R, ..., K] makes it deterministic (the reference draws from the
        global TF RNG)."""
        Fmu, Fvar = np.asarray(Fmu, dtype=DEFAULT_FLOAT), np.asarray(Fvar, dtype=DEFAULT_FLOAT)
        if eps is None:
            eps = (rng or np.random.default_rng()).standard_normal((self.num_monte_carlo_points,) + Fmu.shape)
        p = self.conditional_mean(Fmu[None] + np.sqrt(Fvar)[None] * np.asarray(eps, dtype=DEFAULT_FLOAT))
        mean = p.mean(0)
        return mean, (p * p).mean(0) - mean * mean


# ---- mean functions ---------------------------------------------------------------------------------------
class MeanFunction(Module):
    pass


class Zero(MeanFunction):
    def __init__(self, output_dim=-1):
        self.output_dim = output_dim


class Constant(MeanFunction):
    def __init__(self, c=None):
        self.c = Parameter(np.zeros(1) if c is None else c, name="c")


# ---- inducing variables ---------------------------------------------------------------------------------------
class InducingVariables(Module):
    pass


class InducingPoints(InducingVariables):
    def __init__(self, Z, name=None):
        Z = np.asarray(Z, dtype=DEFAULT_FLOAT)
        if Z.ndim != 2:
            raise ValueError("Z must be [M, D]")
        self.Z = Parameter(Z, name="Z")
        self.name = name

    @property
    def num_inducing(self):
        return self.Z.unconstrained.shape[0]


class SharedIndependentInducingVariables(InducingVariables):
    def __init__(self, inducing_variable: InducingPoints):
        self.inducing_variable = inducing_variable

    @property
    def Z(self):
        return self.inducing_variable.Z

    @property
    def num_inducing(self):
        return self.inducing_variable.num_inducing


def inducingpoint_wrapper(iv):
    if isinstance(iv, InducingVariables):
        return iv
    return InducingPoints(iv)


# ---- SVGP parameter container ------------------------------------------------------------------------------
class SVGP(Module):
    """Parameter container with gpflow.models.SVGP's constructor (whiten=True, q_diag=False)."""

    def __init__(self, kernel, likelihood, inducing_variable, mean_function=None, num_latent_gps: int = 1,
                 q_diag: bool = False, q_mu=None, q_sqrt=None, whiten: bool = True, num_data=None):
        if not whiten:
            raise NotImplementedError("whiten=False is not implemented by the CUDA path (SURVEY 8f-4)")
        self.kernel = kernel
        self.likelihood = likelihood
        self.inducing_variable = inducingpoint_wrapper(inducing_variable)
        self.mean_function = mean_function if mean_function is not None else Zero()
        self.num_latent_gps = int(num_latent_gps)
        self.q_diag, self.whiten, self.num_data = q_diag, whiten, num_data
        M, L = self.inducing_variable.num_inducing, self.num_latent_gps
        q_mu = np.zeros((M, L)) if q_mu is None else np.asarray(q_mu, dtype=DEFAULT_FLOAT)
        if q_mu.shape != (M, L):
            raise ValueError(f"q_mu must be [{M}, {L}], got {q_mu.shape}")
        self.q_mu = Parameter(q_mu, name="q_mu")
        if q_diag:
            # gpflow SVGP._init_variational_parameters: a diagonal covariance, q_sqrt [M, L] with the positive() transform.  The
            # CUDA path sees L diagonal [M, M] factors (q_sqrt_full): the same conditional, KL and gradients, restricted to the
            # diagonal entries -- the only variables of those matrices.
            q_sqrt = np.ones((M, L)) if q_sqrt is None else np.asarray(q_sqrt, dtype=DEFAULT_FLOAT)
            if q_sqrt.shape != (M, L):
                raise ValueError(f"q_sqrt must be [{M}, {L}] for q_diag=True, got {q_sqrt.shape}")
            self.q_sqrt = Parameter(q_sqrt, transform=positive(), name="q_sqrt")
        else:
            q_sqrt = np.tile(np.eye(M)[None], (L, 1, 1)) if q_sqrt is None else np.asarray(q_sqrt, dtype=DEFAULT_FLOAT)
            if q_sqrt.shape != (L, M, M):
                raise ValueError(f"q_sqrt must be [{L}, {M}, {M}], got {q_sqrt.shape}")
            self.q_sqrt = Parameter(np.tril(q_sqrt), transform=triangular(), name="q_sqrt")
        self._kernels = latent_kernels(kernel, L)

    @staticmethod
    def diag_embed(d):
        """[M, L] -> [L, M, M] with d[:, l] on the diagonal of matrix l."""
        d = np.asarray(d)
        out = np.zeros((d.shape[1], d.shape[0], d.shape[0]), dtype=d.dtype)
        idx = np.arange(d.shape[0])
        out[:, idx, idx] = d.T
        return out

    def q_sqrt_full(self, values=None):
        """The [L, M, M] Cholesky factors the CUDA path works with (`values`: a stand-in for the q_sqrt entries, e.g. variable ids)."""
        v = self.q_sqrt.numpy() if values is None else values
        return self.diag_embed(v) if self.q_diag else v

    @property
    def Z(self) -> Parameter:
        return self.inducing_variable.Z

    @property
    def latent_kernel_list(self):
        return self._kernels
