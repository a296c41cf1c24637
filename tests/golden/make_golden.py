#!/usr/bin/env python3
"""Generates tests/golden/golden_v1.npz.

The reference cannot run in this environment (TensorFlow / GPflow-fork / TFP are not installable and the
snapshot lacks mogpe/custom_types.py), so these vectors are produced by the CPU oracle
(oracle/ref_numpy.py for values, oracle/ref_torch.py for gradients) -- PARITY UNPINNED by the reference.
They freeze the oracle's answers on (a) the reference's own test-fixture geometry
(tests/keras/test_keras.py:16-49), (b) the reference's mcycle data set (examples/mcycle/data/mcycle.csv,
standardised as examples/mcycle/data/load_data.py:24-27 does) with the parameters of
examples/mcycle/configs/config_2_experts.toml:21-81, and (c) seeded random models for every flavour x bound.
The GPU tests compare the CUDA path against these numbers WITHOUT running the oracle.

Run from the repo root in the build container (needs /root/reference for the mcycle csv):
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_numpy as rn  # noqa: E402
from oracle import ref_torch as rt  # noqa: E402
from tests._specs import BOUNDS, FLAVOURS, make_spec  # noqa: E402

MCYCLE = "/root/reference/examples/mcycle/data/mcycle.csv"


def mcycle_case():
    data = np.loadtxt(MCYCLE, delimiter=",", skiprows=1, usecols=(1, 2))
    X = data[:, 0].reshape(-1, 1)
    Y = data[:, 1].reshape(-1, 1)
    X = (X - X.mean()) / X.std()
    Y = (Y - Y.mean()) / Y.std()
    rng = np.random.default_rng(0)
    M = 32

    def gp(ls, var, noise, q_sqrt_scale, c):
        idx = rng.choice(X.shape[0], M, replace=False)
        d = {"Z": X[idx].copy(), "kernels": [{"lengthscales": ls, "variance": var}],
             "q_mu": 2.0 * rng.standard_normal((M, 1)), "q_sqrt": q_sqrt_scale * np.eye(M)[None], "mean_c": c}
        if noise is not None:
            d["noise_variance"] = np.array([noise])
        return d

    experts = [gp(10.0, 0.1, 0.032, 1.0, np.array([0.0])), gp(0.5, 20.0, 0.9, 1.0, np.array([0.0]))]
    gating = gp(0.5, 3.0, None, 2.0, None)
    return X, Y, experts, gating


def main():
    out = {}
    # (a) reference fixture geometry
    N, M = 5, 4
    kern = {"lengthscales": np.array([3.0, 3.0]), "variance": 10.0}
    gp = {"Z": np.ones((M, 2)), "kernels": [kern], "q_mu": np.zeros((M, 1)), "q_sqrt": np.eye(M)[None].copy(),
          "mean_c": np.array([-2.0]), "noise_variance": np.array([3.0])}
    gate = {k: v for k, v in gp.items() if k != "noise_variance"}
    X, Y = np.ones((N, 2)), np.ones((N, 1))
    eps = {"u_experts": [np.full((1, 1, M), 0.3), np.full((1, 1, M), -0.1)], "h": np.full((1, N, 1), -0.2)}
    for fl in FLAVOURS:
        spec = {"flavour": fl, "bound": "further_gating", "num_samples": 1, "num_data": None, "experts": [gp, gp], "gating": gate}
        out[f"fixture/{fl}/elbo"] = rn.elbo(spec, X, Y, eps)
        m, v = rn.predict_y(spec, X)
        out[f"fixture/{fl}/y_mean"], out[f"fixture/{fl}/y_var"] = m, v
    # (b) mcycle with the TOML parameters, legacy flavour (the TOML front-end builds legacy models)
    X, Y, experts, gating = mcycle_case()
    out["mcycle/X"], out["mcycle/Y"] = X, Y
    for k, g in enumerate(experts):
        out[f"mcycle/expert{k}/Z"], out[f"mcycle/expert{k}/q_mu"] = g["Z"], g["q_mu"]
    out["mcycle/gating/Z"], out["mcycle/gating/q_mu"] = gating["Z"], gating["q_mu"]
    rng = np.random.default_rng(1)
    idx = rng.choice(X.shape[0], 16, replace=False)
    out["mcycle/batch_idx"] = idx
    eps = {"u_experts": [rng.standard_normal((1, 1, 32)) for _ in range(2)], "u_gating": rng.standard_normal((1, 1, 32)),
           "h": rng.standard_normal((1, 16, 1)), "f": rng.standard_normal((1, 16, 1, 2))}
    for k in ("u_gating", "h", "f"):
        out[f"mcycle/eps/{k}"] = eps[k]
    out["mcycle/eps/u_experts"] = np.stack(eps["u_experts"])
    for fl in FLAVOURS:
        for b in BOUNDS:
            spec = {"flavour": fl, "bound": b, "num_samples": 1, "num_data": X.shape[0], "experts": experts, "gating": gating}
            out[f"mcycle/{fl}/{b}/elbo"] = rn.elbo(spec, X[idx], Y[idx], eps)
        spec = {"flavour": fl, "bound": "further_gating", "num_samples": 1, "num_data": X.shape[0], "experts": experts, "gating": gating}
        m, v = rn.predict_y(spec, X)
        out[f"mcycle/{fl}/y_mean"], out[f"mcycle/{fl}/y_var"] = m, v
        out[f"mcycle/{fl}/mix_probs"] = rn.predict_mixing_probs(spec, X)
        _, grads = rt.MixtureOfSVGPExperts(spec).elbo_and_grads(X[idx], Y[idx], eps)
        for name in ("experts/0/kernels/0/lengthscales", "experts/1/noise_variance", "gating/q_mu", "gating/Z", "experts/1/q_sqrt"):
            out[f"mcycle/{fl}/grad/{name}"] = grads[name]
    # (c) seeded random models: every flavour x bound, K = 2 and K = 3
    for fl in FLAVOURS:
        for b in BOUNDS:
            for K in (2, 3):
                if b == "tight" and K > 2:
                    continue
                spec, X, Y, eps = make_spec(seed=1000 + K, N=64, D=2, F=2, K=K, M=12, S=2, flavour=fl, bound=b,
                                            num_data=6400, separate_kernels=True, gating_mean_c=0.1)
                out[f"random/{fl}/{b}/K{K}/elbo"] = rn.elbo(spec, X, Y, eps)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), **out)
    print(f"wrote {len(out)} arrays")


if __name__ == "__main__":
    main()
