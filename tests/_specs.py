"""Seeded model specs + explicit normal draws shared by the oracle tests and the GPU parity tests.

The spec format is documented in oracle/ref_numpy.py.  Parameter distributions follow SURVEY.md 8(d):
lengthscales, kernel variance ~ LogU(0.3, 3), noise ~ LogU(1e-3, 1), q_mu ~ N(0, 1),
q_sqrt = tril(0.1 N(0,1)) + I, Z = M random data rows (mogpe/keras/utils.py:83-93).
"""
import numpy as np


def logu(rng, lo, hi, size=None):
    return np.exp(rng.uniform(np.log(lo), np.log(hi), size))


def make_gp(rng, X, M, L, n_kernels, ard, mean_c, is_expert, noise_lo=1e-2, jitter_z=0.0, ls_range=(0.3, 3.0)):
    N, D = X.shape
    idx = rng.choice(N, size=M, replace=N < M)
    Z = X[idx].copy() + jitter_z * rng.standard_normal((M, D))
    kernels = []
    for _ in range(n_kernels):
        ls = logu(rng, ls_range[0], ls_range[1], D) if ard else float(logu(rng, ls_range[0], ls_range[1]))
        kernels.append({"lengthscales": ls, "variance": float(logu(rng, 0.3, 3.0))})
    gp = {
        "Z": Z,
        "kernels": kernels,
        "q_mu": rng.standard_normal((M, L)),
        "q_sqrt": np.tril(0.1 * rng.standard_normal((L, M, M))) + np.eye(M)[None],
        "mean_c": None if mean_c is None else np.full((L,), float(mean_c)) + 0.1 * rng.standard_normal(L),
    }
    if is_expert:
        gp["noise_variance"] = logu(rng, noise_lo, 1.0, L)
    return gp


def make_data(rng, N, D, F, K):
    X = rng.standard_normal((N, D))
    regime = rng.integers(0, K, N)
    Y = np.stack([np.sin((1 + f) * X.sum(-1) + regime) + 0.1 * (1 + regime) * rng.standard_normal(N) for f in range(F)], -1)
    return X, Y


def make_spec(seed=0, N=24, D=2, F=1, K=2, M=8, S=1, flavour="keras", bound="further_gating",
              num_data=None, separate_kernels=False, ard=True, mean_c=0.3, gating_mean_c=None,
              M_gating=None, M_experts=None, noise_lo=None, ls_range=(0.3, 3.0)):
    """-> spec, X, Y, eps (dict of explicit standard-normal draws)."""
    rng = np.random.default_rng(seed)
    X, Y = make_data(rng, N, D, F, K)
    G = 1 if K == 2 else K
    Ms = M_experts if M_experts is not None else [M] * K
    if noise_lo is None:
        # "further" uses the noise *variance* as the density scale (quirk 2): keep it large enough that the
        # reference's probability-space evaluation (quirk 6) stays finite on these seeds
        # (the keras flavour also uses y_var itself as the scale of the expert density, quirk 1)
        noise_lo = 0.5 if bound == "further" else (0.3 if flavour == "keras" else 1e-2)
    experts = [
        make_gp(rng, X, Ms[k], F, F if separate_kernels else 1, ard, mean_c, True, noise_lo=noise_lo, ls_range=ls_range)
        for k in range(K)
    ]
    gating = make_gp(rng, X, M_gating or M, G, G if (separate_kernels and G > 1) else 1, ard, gating_mean_c, False,
                     ls_range=ls_range)
    spec = {"flavour": flavour, "bound": bound, "num_samples": S, "num_data": num_data,
            "experts": experts, "gating": gating}
    eps = {
        "u_experts": [rng.standard_normal((S, F, Ms[k])) for k in range(K)],
        "u_gating": rng.standard_normal((S, G, M_gating or M)),
        "h": rng.standard_normal((S, N, G)),
        "f": rng.standard_normal((S, N, F, K)),
    }
    return spec, X, Y, eps


FLAVOURS = ["keras", "legacy"]
BOUNDS = ["further_gating", "tight", "further"]


def restrict_active_dims(spec, expert_dims, gating_dims):
    """Give every kernel of the spec gpflow `active_dims` (column indices) and cut its ARD lengthscales down to them."""
    def cut(gp, dims):
        for k in gp["kernels"]:
            k["active_dims"] = list(dims)
            ls = np.asarray(k["lengthscales"], dtype=float)
            if ls.ndim:
                k["lengthscales"] = ls[list(dims)].copy()
    for gp in spec["experts"]:
        cut(gp, expert_dims)
    cut(spec["gating"], gating_dims)
    return spec


def make_q_diag(spec, experts=True, gating=True):
    """Turn the variational covariances of the spec into gpflow's q_diag=True form: q_sqrt [M, L] = the (positive) diagonals."""
    for gp in (spec["experts"] if experts else []) + ([spec["gating"]] if gating else []):
        q = np.asarray(gp["q_sqrt"])
        gp["q_sqrt"] = np.abs(np.diagonal(q, axis1=-2, axis2=-1)).T.copy()
        gp["q_diag"] = True
    return spec
