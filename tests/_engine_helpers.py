"""Bridges between the oracle's spec dicts and the product's Engine (tests only)."""
import numpy as np

from mogpe_b200.engine import Engine, GPBlock


def rn_q_sqrt_full(gp):
    from oracle.ref_numpy import q_sqrt_full

    return q_sqrt_full(gp)


def block_from_gp(gp, is_expert):
    L = np.asarray(gp["q_mu"]).shape[1]
    return GPBlock(
        Z=np.asarray(gp["Z"], dtype=np.float64),
        lengthscales=[np.asarray(k["lengthscales"], dtype=np.float64) for k in gp["kernels"]],
        variances=[float(k["variance"]) for k in gp["kernels"]],
        q_mu=np.asarray(gp["q_mu"], dtype=np.float64),
        q_sqrt=rn_q_sqrt_full(gp),
        mean_c=None if gp.get("mean_c") is None else np.asarray(gp["mean_c"], dtype=np.float64),
        noise=np.broadcast_to(np.asarray(gp["noise_variance"], dtype=np.float64).reshape(-1), (L,)).copy() if is_expert else None,
        active_dims=([k.get("active_dims") for k in gp["kernels"]] if any(k.get("active_dims") is not None for k in gp["kernels"])
                     else None),
    )


def engine_from_spec(spec, max_batch, precision="float64"):
    experts = [block_from_gp(g, True) for g in spec["experts"]]
    gating = block_from_gp(spec["gating"], False)
    return Engine(experts, gating, flavour=spec["flavour"], max_batch=max_batch, precision=precision)


def run_engine_elbo(eng, spec, X, Y, eps, want_grads=True, device_inputs=False):
    S = spec["num_samples"]
    b = spec["bound"]
    N = X.shape[0]
    scale = 1.0 if spec.get("num_data") is None else spec["num_data"] / N
    eps_u = eps_h = eps_f = None
    if b == "further_gating":
        eps_u = eng.pack_eps_u(eps["u_experts"], None, S)
        eps_h = np.ascontiguousarray(eps["h"])
    elif b == "tight":
        eps_u = eng.pack_eps_u(eps["u_experts"], eps["u_gating"], S)
        if eng.G > 1:
            eng.set_mc(eps["mc"])  # Softmax likelihood expectation is Monte-Carlo: [S, R, N, G]
    else:
        eps_h = np.ascontiguousarray(eps["h"])
        eps_f = eng.pack_eps_f(eps["f"])
    if device_inputs:
        import torch

        dev = lambda a: None if a is None else torch.as_tensor(a, device="cuda")
        return eng.elbo(dev(X), dev(Y), b, S, scale, dev(eps_u), dev(eps_h), dev(eps_f), want_grads=want_grads)
    return eng.elbo(X, Y, b, S, scale, eps_u, eps_h, eps_f, want_grads=want_grads)


def oracle_grads_constrained(spec, grads_u):
    """torch-oracle gradients (w.r.t. unconstrained variables) -> constrained space, as
    {name: array} with the oracle's names."""
    from oracle import ref_numpy as rn

    out = {}
    for name, g in grads_u.items():
        parts = name.split("/")
        gp = spec["gating"] if parts[0] == "gating" else spec["experts"][int(parts[1])]
        key = "/".join(parts[1:]) if parts[0] == "gating" else "/".join(parts[2:])
        g = np.asarray(g)
        if key in ("Z", "q_mu", "mean_c"):
            out[name] = g
        elif key == "q_sqrt":
            M = np.asarray(gp["q_sqrt"]).shape[-1]
            idx = rn.fill_triangular_index(M)
            out[name] = np.tril(g[..., idx])
        elif key == "noise_variance":
            v = np.asarray(gp["noise_variance"], dtype=float).reshape(-1)
            out[name] = g / (-np.expm1(-(v - 1e-6)))
        else:  # kernels/i/lengthscales|variance
            ki, field = int(key.split("/")[1]), key.split("/")[2]
            v = np.asarray(gp["kernels"][ki][field], dtype=float)
            out[name] = g / (-np.expm1(-v))
    return out


def engine_grads_named(res, spec):
    """ElboResult -> {oracle name: array} (constrained space)."""
    out = {}
    blocks = [(f"experts/{k}", g, spec["experts"][k]) for k, g in enumerate(res.experts)] + [("gating", res.gating, spec["gating"])]
    for prefix, g, gp in blocks:
        out[f"{prefix}/Z"] = g.Z
        for i in range(len(g.lengthscales)):
            out[f"{prefix}/kernels/{i}/lengthscales"] = np.asarray(g.lengthscales[i])
            out[f"{prefix}/kernels/{i}/variance"] = np.asarray(g.variances[i])
        out[f"{prefix}/q_mu"] = g.q_mu
        out[f"{prefix}/q_sqrt"] = g.q_sqrt
        if gp.get("mean_c") is not None:
            out[f"{prefix}/mean_c"] = g.mean_c
        if g.noise is not None:
            out[f"{prefix}/noise_variance"] = g.noise
    return out
