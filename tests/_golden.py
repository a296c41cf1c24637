"""Rebuild the model specs stored in tests/golden/golden_v1.npz (see tests/golden/make_golden.py)."""
import os

import numpy as np

from tests._specs import make_spec

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


def load():
    return np.load(PATH)


def fixture_case(flavour):
    N, M = 5, 4
    kern = {"lengthscales": np.array([3.0, 3.0]), "variance": 10.0}
    gp = {"Z": np.ones((M, 2)), "kernels": [kern], "q_mu": np.zeros((M, 1)), "q_sqrt": np.eye(M)[None].copy(),
          "mean_c": np.array([-2.0]), "noise_variance": np.array([3.0])}
    gate = {k: v for k, v in gp.items() if k != "noise_variance"}
    spec = {"flavour": flavour, "bound": "further_gating", "num_samples": 1, "num_data": None, "experts": [gp, gp], "gating": gate}
    eps = {"u_experts": [np.full((1, 1, M), 0.3), np.full((1, 1, M), -0.1)], "h": np.full((1, N, 1), -0.2)}
    return spec, np.ones((N, 2)), np.ones((N, 1)), eps


def mcycle_case(g, flavour, bound):
    """mcycle data + config_2_experts.toml parameters (reference examples/mcycle/configs/config_2_experts.toml:21-81)."""
    M = 32

    def gp(prefix, ls, var, noise, q_sqrt_scale, c):
        d = {"Z": g[f"mcycle/{prefix}/Z"], "kernels": [{"lengthscales": ls, "variance": var}], "q_mu": g[f"mcycle/{prefix}/q_mu"],
             "q_sqrt": q_sqrt_scale * np.eye(M)[None], "mean_c": c}
        if noise is not None:
            d["noise_variance"] = np.array([noise])
        return d

    experts = [gp("expert0", 10.0, 0.1, 0.032, 1.0, np.array([0.0])), gp("expert1", 0.5, 20.0, 0.9, 1.0, np.array([0.0]))]
    gating = gp("gating", 0.5, 3.0, None, 2.0, None)
    X, Y = g["mcycle/X"], g["mcycle/Y"]
    spec = {"flavour": flavour, "bound": bound, "num_samples": 1, "num_data": X.shape[0], "experts": experts, "gating": gating}
    eps = {"u_experts": list(g["mcycle/eps/u_experts"]), "u_gating": g["mcycle/eps/u_gating"], "h": g["mcycle/eps/h"],
           "f": g["mcycle/eps/f"]}
    return spec, X, Y, g["mcycle/batch_idx"], eps


def random_case(flavour, bound, K):
    return make_spec(seed=1000 + K, N=64, D=2, F=2, K=K, M=12, S=2, flavour=flavour, bound=bound, num_data=6400,
                     separate_kernels=True, gating_mean_c=0.1)
