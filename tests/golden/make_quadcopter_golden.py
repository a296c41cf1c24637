#!/usr/bin/env python3
"""Generates tests/golden/quadcopter_golden.npz from the reference's own trained checkpoint
(examples/quadcopter/saved_ckpts/.../11-04-210440/ckpt-200, copied as DATA to tests/golden/quadcopter_ckpt/) and a
256-row subset of the reference's quadcopter data (examples/quadcopter/data/quadcopter_data_step_10_direction_down.npz,
first two outputs, standardised as examples/quadcopter/data/load_data.py:44-46).

The parameters are reference-TRAINED (2 experts x 2 outputs with separate ARD kernels, M = 100, Bernoulli gating; one
pre-softplus lengthscale is 97.6, noise variances down to 6e-3); the outputs frozen here are the CPU oracle's (legacy
flavour, as the checkpoint was written by the legacy training loop) -- the reference itself cannot run in this image.
Run from the repo root in the build container:  python tests/golden/make_quadcopter_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mogpe_b200.training.tf_checkpoint import read_tf_checkpoint  # noqa: E402  (pure-python table reader)
from oracle import ref_numpy as rn  # noqa: E402

DATA = "/root/reference/examples/quadcopter/data/quadcopter_data_step_10_direction_down.npz"


def spec_from_tensors(t, bound):
    def g(prefix, expert):
        def v(k):
            return t[f"{prefix}/{k}/_pretransformed_input/.ATTRIBUTES/VARIABLE_VALUE"]

        zkey = "inducing_variable/inducing_variable/Z" if expert else "inducing_variable/Z"
        nk = 2 if expert else 1
        d = {"Z": v(zkey),
             "kernels": [{"lengthscales": rn.softplus(v(f"kernel/kernels/{i}/lengthscales")),
                          "variance": float(rn.softplus(v(f"kernel/kernels/{i}/variance"))[0])} for i in range(nk)],
             "q_mu": v("q_mu"), "q_sqrt": rn.fill_triangular(v("q_sqrt")), "mean_c": v("mean_function/c") if expert else None}
        if expert:
            d["noise_variance"] = (1e-6 + rn.softplus(v("likelihood/variance"))).reshape(-1)
        return d

    return {"flavour": "legacy", "bound": bound, "num_samples": 1, "num_data": 2685,
            "experts": [g(f"model/experts/experts_list/{k}", True) for k in range(2)], "gating": g("model/gating_network", False)}


def main():
    t = read_tf_checkpoint(os.path.join(ROOT, "tests", "golden", "quadcopter_ckpt", "ckpt-200"))
    data = np.load(DATA)
    X, Y = data["x"], data["y"][:, :2]
    X = (X - X.mean()) / X.std()
    Y = (Y - Y.mean()) / Y.std()
    rng = np.random.default_rng(0)
    idx = np.sort(rng.choice(X.shape[0], 256, replace=False))
    X, Y = np.ascontiguousarray(X[idx]), np.ascontiguousarray(Y[idx])
    eps = {"u_experts": [rng.standard_normal((1, 2, 100)) for _ in range(2)], "u_gating": rng.standard_normal((1, 1, 100)),
           "h": rng.standard_normal((1, 256, 1)), "f": rng.standard_normal((1, 256, 2, 2))}
    out = {"X": X, "Y": Y, "eps/u_experts": np.stack(eps["u_experts"]), "eps/u_gating": eps["u_gating"], "eps/h": eps["h"],
           "eps/f": eps["f"]}
    for b in ("further_gating", "tight", "further"):
        out[f"elbo/{b}"] = rn.elbo(spec_from_tensors(t, b), X, Y, eps)
    spec = spec_from_tensors(t, "further_gating")
    out["y_mean"], out["y_var"] = rn.predict_y(spec, X)
    out["mix_probs"] = rn.predict_mixing_probs(spec, X)
    fm, fv = rn.predict_f(spec["experts"][0], X)
    out["expert0/f_mean"], out["expert0/f_var"] = fm, fv
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "quadcopter_golden.npz"), **out)
    print({k: (float(v) if np.ndim(v) == 0 else np.shape(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
