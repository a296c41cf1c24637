"""Build a legacy-flavour MixtureOfSVGPExperts from the reference's TOML schema
(mogpe/training/toml_config_parsers/model_parsers.py:414-431; example:
examples/mcycle/configs/config_2_experts.toml).  Same keys, same defaults (and the same printed notes when a key is
missing): every field is optional except num_experts.

Schema (all under the top level): num_experts, num_samples, bound, [[experts]] with whiten, mean_function.{name,params.constant},
kernel.{name,params.{lengthscale,variance}} (a table, or a list of tables = one kernel per output),
likelihood.{name,params.variance}, inducing_points.{num_inducing,q_mu.{mean,var},q_sqrt,q_diag,inputs_trainable};
[gating_network] with the same svgp fields (kernel list = one kernel per gating function).
"""
try:
    import tomllib as _toml
except ImportError:  # pragma: no cover
    import toml as _toml

import numpy as np

from ... import gpflow_lite as gl
from ...experts import SVGPExpert, SVGPExperts
from ...gating_networks import SVGPGatingNetwork
from ...mixture_of_experts import MixtureOfSVGPExperts

_BOUNDS = ("further", "tight", "further_gating", "further_expert", "tight_2")


def _get(cfg, path, default=None, note=None):
    cur = cfg
    for key in path.split("."):
        if not isinstance(cur, dict) or key not in cur:
            if note:
                print(note)
            return default
        cur = cur[key]
    return cur


def parse_single_kernel(kernel_cfg, input_dim):
    name = kernel_cfg.get("name")
    if name != "rbf":
        raise NotImplementedError(f"Kernel name {name} is not on the CUDA path (only 'rbf'; the reference also parses "
                                  "'cosine' and 'product')")
    active_dims = _get(kernel_cfg, "active_dims")  # reference parse_active_dims / parse_lengthscale (model_parsers.py:18-44, 287-292)
    n_ls = input_dim if active_dims is None else len(active_dims)
    ls = _get(kernel_cfg, "params.lengthscale", 1.0, "No lengthscale in config so setting it to 1.0")
    var = _get(kernel_cfg, "params.variance", 1.0, "No variance in config so setting it to 1.0")
    # ARD vector (one entry per active dimension), as the reference builds
    return gl.RBF(lengthscales=np.full((n_ls,), float(ls)), variance=float(var),
                  active_dims=None if active_dims is None else [int(d) for d in active_dims])


def parse_kernel(kernels_cfg, input_dim, output_dim):
    if isinstance(kernels_cfg, list):
        if len(kernels_cfg) == 1 and output_dim == 1:
            return parse_single_kernel(kernels_cfg[0], input_dim)
        return gl.SeparateIndependent([parse_single_kernel(k, input_dim) for k in kernels_cfg])
    if output_dim > 1:
        return gl.SeparateIndependent([parse_single_kernel(kernels_cfg, input_dim) for _ in range(output_dim)])
    return parse_single_kernel(kernels_cfg, input_dim)


def parse_likelihood(likelihood_cfg, output_dim):
    if likelihood_cfg is None or likelihood_cfg.get("name") != "gaussian":
        raise NotImplementedError("This likelihood cannot be instantiated using toml config")
    var = _get(likelihood_cfg, "params.variance", None, "No variance found for Gaussian likelihood")
    if var is None:
        return gl.Gaussian()
    # the reference wraps the TOML value in a list (parse_variance, model_parsers.py:30-35): a scalar becomes a [1] variable,
    # a per-output list such as `variance = [0.0011, 0.0011]` (examples/quadcopter/configs/*.toml) a [1, output_dim] one
    v = np.asarray(var, dtype=np.float64)
    if v.ndim == 0:
        return gl.Gaussian(variance=float(v))
    if v.size not in (1, output_dim):
        raise ValueError(f"likelihood.params.variance has {v.size} entries for {output_dim} outputs")
    return gl.Gaussian(variance=v.reshape(1, -1))


def parse_mean_function(svgp_cfg):
    name = _get(svgp_cfg, "mean_function.name", None, "No mean_function.name in toml config so using Zero()")
    if name is None or name == "zero":
        return gl.Zero()
    if name == "constant":
        return gl.Constant(_get(svgp_cfg, "mean_function.params.constant", None,
                                "No mean_function.constant in toml config so using the default"))
    raise NotImplementedError("This mean function cannot be instantiated using toml config (only zero and constant)")


def parse_inducing_points(svgp_cfg, output_dim, rng):
    """-> q_mu, q_sqrt, q_diag, inputs_trainable.  q_mu = mean + N(0,1) * q_mu.var; q_sqrt = tril(q_sqrt * |N(0,1)|)
    (the reference passes the full matrix to a FillTriangular parameter, which keeps the lower triangle)."""
    ip = _get(svgp_cfg, "inducing_points")
    if ip is None:
        print("inducing_points not found in toml config, setting q_mu=None, q_sqrt=None, q_diag=False")
        return None, None, False, True
    M = ip.get("num_inducing")
    q_mu = q_sqrt = None
    mean, var = _get(ip, "q_mu.mean"), _get(ip, "q_mu.var")
    if mean is not None and var is not None and M is not None:
        q_mu = mean * np.ones((M, 1)) + rng.standard_normal((M, output_dim)) * var
    if ip.get("q_sqrt") is not None and M is not None:
        q_sqrt = np.array([np.tril(ip["q_sqrt"] * np.eye(M) @ np.abs(rng.standard_normal((M, M)))) for _ in range(output_dim)])
    q_diag = bool(ip.get("q_diag", False))
    if q_diag and q_sqrt is not None:
        q_sqrt = np.diagonal(q_sqrt, axis1=-2, axis2=-1).T.copy()  # gpflow's q_diag=True parameter is [M, L]
    return q_mu, q_sqrt, q_diag, bool(ip.get("inputs_trainable", True))


def parse_inducing_variable(svgp_cfg, X, rng):
    M = _get(svgp_cfg, "inducing_points.num_inducing", None, "num_inducing not found in toml config so setting as num_data/4")
    M = int(X.shape[0] / 4) if M is None else int(M)
    idx = rng.choice(X.shape[0], size=M, replace=False)
    return gl.InducingPoints(np.asarray(X)[idx, :].reshape(M, X.shape[1]))


def parse_expert(expert_cfg, input_dim, output_dim, X, rng):
    q_mu, q_sqrt, q_diag, inputs_trainable = parse_inducing_points(expert_cfg, output_dim, rng)
    iv = parse_inducing_variable(expert_cfg, X, rng)
    iv.Z.trainable = inputs_trainable
    return SVGPExpert(parse_kernel(expert_cfg.get("kernel"), input_dim, output_dim),
                      parse_likelihood(expert_cfg.get("likelihood"), output_dim), iv,
                      mean_function=parse_mean_function(expert_cfg), num_latent_gps=output_dim, q_diag=q_diag, q_mu=q_mu,
                      q_sqrt=q_sqrt, whiten=bool(expert_cfg.get("whiten", True)), num_data=X.shape[0])


def parse_gating_network(cfg, X, rng):
    g = cfg.get("gating_network")
    if g is None:
        raise NotImplementedError("gating_network not specified in toml config")
    if g.get("active_dims") is not None:
        raise NotImplementedError("gating_network.active_dims (the reference slices the gating inducing inputs but not the "
                                  "gating kernel's view of X, model_parsers.py:331-336) is not supported: put active_dims on "
                                  "the gating kernel instead")
    if "num_experts" not in cfg:
        raise NotImplementedError("num_experts not specified in toml config")
    K = int(cfg["num_experts"])
    G = K if K > 2 else 1
    lik = gl.Softmax(K) if K > 2 else gl.Bernoulli()
    q_mu, q_sqrt, q_diag, inputs_trainable = parse_inducing_points(g, G, rng)
    iv = parse_inducing_variable(g, X, rng)
    iv.Z.trainable = inputs_trainable
    return SVGPGatingNetwork(parse_kernel(g.get("kernel"), X.shape[1], G), lik, iv, parse_mean_function(g),
                             num_gating_functions=G, q_diag=q_diag, q_mu=q_mu, q_sqrt=q_sqrt,
                             whiten=bool(g.get("whiten", True)), num_data=X.shape[0])


def parse_bound(cfg):
    b = cfg.get("bound")
    if b is None:
        print("No bound in toml config so setting bound='further'")
        return "further"
    if b not in _BOUNDS:
        print("Bound in toml config should be either 'further' or 'tight',  setting bound='further'")
        return "further"
    return b


def MixtureOfSVGPExperts_from_toml(config_file, dataset, seed=None):
    """(config_file, (X, Y)) -> legacy MixtureOfSVGPExperts.  `seed` makes the random parts of the initialisation
    (Z rows, q_mu noise, q_sqrt) reproducible; the reference draws them from numpy's global RNG."""
    with open(config_file, "rb") as f:
        cfg = _toml.load(f) if hasattr(_toml, "load") and _toml.__name__ == "tomllib" else _toml.loads(f.read().decode())
    rng = np.random.default_rng(seed)
    X, Y = (np.asarray(a, dtype=np.float64) for a in dataset)
    num_data, input_dim = X.shape
    output_dim = Y.shape[1]
    experts = SVGPExperts([parse_expert(e, input_dim, output_dim, X, rng) for e in cfg.get("experts", [])])
    gating = parse_gating_network(cfg, X, rng)
    return MixtureOfSVGPExperts(gating_network=gating, experts=experts, num_samples=int(cfg.get("num_samples", 1)),
                                num_data=num_data, bound=parse_bound(cfg))
