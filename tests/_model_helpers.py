"""Build product models (both API flavours) from the oracle's spec dicts (tests only)."""
import numpy as np

from mogpe_b200 import gpflow_lite as gl


def _kernel(gp, L):
    ks = [gl.RBF(variance=k["variance"], lengthscales=np.asarray(k["lengthscales"], dtype=float), active_dims=k.get("active_dims"))
          for k in gp["kernels"]]
    if len(ks) == 1:
        return ks[0] if L == 1 else gl.SharedIndependent(ks[0], L)
    return gl.SeparateIndependent(ks)


def _mean(gp):
    return None if gp.get("mean_c") is None else gl.Constant(np.asarray(gp["mean_c"], dtype=float))


def keras_model_from_spec(spec):
    from mogpe_b200.keras import MixtureOfSVGPExperts, SVGPExpert, SVGPGatingNetwork

    experts = []
    for gp in spec["experts"]:
        L = np.asarray(gp["q_mu"]).shape[1]
        kern = _kernel(gp, L)
        if isinstance(kern, gl.SharedIndependent):
            kern = kern.kernel
        experts.append(SVGPExpert(kern, gl.Gaussian(np.asarray(gp["noise_variance"], dtype=float).reshape(1, -1)),
                                  gl.InducingPoints(gp["Z"]), _mean(gp), num_latent_gps=L, q_mu=gp["q_mu"], q_sqrt=gp["q_sqrt"],
                                  q_diag=bool(gp.get("q_diag"))))
    g = spec["gating"]
    G = np.asarray(g["q_mu"]).shape[1]
    gk = _kernel(g, G)
    gating = SVGPGatingNetwork(gk, gl.InducingPoints(g["Z"]), _mean(g), q_mu=g["q_mu"], q_sqrt=g["q_sqrt"],
                               q_diag=bool(g.get("q_diag")))
    return MixtureOfSVGPExperts(experts, gating, num_samples=spec["num_samples"], num_data=spec.get("num_data"),
                                bound=spec["bound"])


def legacy_model_from_spec(spec):
    from mogpe_b200.experts import SVGPExpert, SVGPExperts
    from mogpe_b200.gating_networks import SVGPGatingNetwork
    from mogpe_b200.mixture_of_experts import MixtureOfSVGPExperts

    experts = []
    for gp in spec["experts"]:
        L = np.asarray(gp["q_mu"]).shape[1]
        kern = _kernel(gp, L)
        if isinstance(kern, gl.SharedIndependent):
            kern = kern.kernel
        experts.append(SVGPExpert(kern, gl.Gaussian(np.asarray(gp["noise_variance"], dtype=float).reshape(1, -1)),
                                  gl.InducingPoints(gp["Z"]), _mean(gp), num_latent_gps=L, q_mu=gp["q_mu"], q_sqrt=gp["q_sqrt"],
                                  q_diag=bool(gp.get("q_diag"))))
    g = spec["gating"]
    G = np.asarray(g["q_mu"]).shape[1]
    gk = _kernel(g, G)
    if isinstance(gk, gl.SharedIndependent):
        gk = gk.kernel if G == 1 else gk
    lik = gl.Bernoulli() if G == 1 else gl.Softmax(G)
    gating = SVGPGatingNetwork(gk, lik, gl.InducingPoints(g["Z"]), _mean(g), num_gating_functions=G, q_mu=g["q_mu"],
                               q_sqrt=g["q_sqrt"], q_diag=bool(g.get("q_diag")))
    return MixtureOfSVGPExperts(gating, SVGPExperts(experts), num_data=spec.get("num_data"),
                                num_samples=spec["num_samples"], bound=spec["bound"])


def model_from_spec(spec):
    return keras_model_from_spec(spec) if spec["flavour"] == "keras" else legacy_model_from_spec(spec)


def oracle_grads_in_model_order(spec, model, grads_u):
    """Map the torch oracle's named unconstrained gradients onto model.trainable_variables (same objects)."""
    comps = (list(model.experts_list) if spec["flavour"] == "keras" else list(model.experts.experts_list)) + [model.gating_network]
    out = {}
    for ci, comp in enumerate(comps):
        gp = comp.gp if hasattr(comp, "gp") else comp
        prefix = f"experts/{ci}" if ci < len(comps) - 1 else "gating"
        out[id(gp.Z)] = grads_u[f"{prefix}/Z"]
        for i, k in enumerate(gp.latent_kernel_list):
            out[id(k.lengthscales)] = grads_u[f"{prefix}/kernels/{i}/lengthscales"]
            out[id(k.variance)] = grads_u[f"{prefix}/kernels/{i}/variance"]
        out[id(gp.q_mu)] = grads_u[f"{prefix}/q_mu"]
        out[id(gp.q_sqrt)] = grads_u[f"{prefix}/q_sqrt"]
        if f"{prefix}/mean_c" in grads_u:
            out[id(gp.mean_function.c)] = grads_u[f"{prefix}/mean_c"]
        if f"{prefix}/noise_variance" in grads_u:
            out[id(gp.likelihood.variance)] = grads_u[f"{prefix}/noise_variance"]
    return [out[id(p)] for p in model.trainable_variables]
