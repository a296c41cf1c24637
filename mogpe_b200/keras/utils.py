"""mogpe/keras/utils.py: Keras-style (de)serialisation of the Keras-flavour model (nested {class_name, config} dicts, as
tf.keras.layers.serialize produces and the reference's YAML files contain) and inducing-input resampling."""
import json
from typing import List, Optional

import numpy as np
import yaml

from .. import gpflow_lite as gl
from .experts import SVGPExpert
from .gating_networks import SVGPGatingNetwork
from .mixture_of_experts import MixtureOfSVGPExperts


class NumpyArrayEncoder(json.JSONEncoder):
    def default(self, obj):
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        if isinstance(obj, np.generic):
            return obj.item()
        return json.JSONEncoder.default(self, obj)


def _arr(v):
    return None if v is None else np.asarray(v, dtype=np.float64)


def _deserialize(cfg, rng):
    """One {class_name, config} node -> object."""
    if cfg is None:
        return None
    name, c = cfg["class_name"], cfg.get("config", {}) or {}
    if name in ("SquaredExponential", "RBF"):
        return gl.RBF(variance=c.get("variance", 1.0), lengthscales=_arr(c.get("lengthscales", 1.0)))
    if name == "SeparateIndependent":
        return gl.SeparateIndependent([_deserialize(k, rng) for k in c["kernels"]])
    if name == "SharedIndependent":
        return gl.SharedIndependent(_deserialize(c["kernel"], rng), c["output_dim"])
    if name == "Gaussian":
        return gl.Gaussian(variance=_arr(c.get("variance", 1.0)), variance_lower_bound=c.get("variance_lower_bound", 1e-6))
    if name == "Constant":
        return gl.Constant(_arr(c.get("c")))
    if name == "Zero":
        return gl.Zero(c.get("output_dim", -1))
    if name == "InducingPoints":
        if c.get("Z") is not None:
            return gl.InducingPoints(_arr(c["Z"]))
        # {num_inducing, input_dim}: placeholder inputs; call sample_mosvgpe_inducing_inputs_from_data(X, model) next,
        # as the reference's notebooks do
        return gl.InducingPoints(rng.standard_normal((int(c["num_inducing"]), int(c["input_dim"]))))
    if name == "SharedIndependentInducingVariables":
        return gl.SharedIndependentInducingVariables(_deserialize(c["inducing_variable"], rng))
    if name == "SVGPExpert":
        return SVGPExpert(kernel=_deserialize(c["kernel"], rng), likelihood=_deserialize(c["likelihood"], rng),
                          inducing_variable=_deserialize(c["inducing_variable"], rng),
                          mean_function=_deserialize(c.get("mean_function"), rng), num_latent_gps=c.get("num_latent_gps") or 1,
                          q_diag=bool(c.get("q_diag") or False), q_mu=_arr(c.get("q_mu")), q_sqrt=_arr(c.get("q_sqrt")),
                          whiten=True if c.get("whiten") is None else bool(c["whiten"]))
    if name == "SVGPGatingNetwork":
        return SVGPGatingNetwork(kernel=_deserialize(c["kernel"], rng), inducing_variable=_deserialize(c["inducing_variable"], rng),
                                 mean_function=_deserialize(c.get("mean_function"), rng), q_diag=bool(c.get("q_diag") or False),
                                 q_mu=_arr(c.get("q_mu")), q_sqrt=_arr(c.get("q_sqrt")),
                                 whiten=True if c.get("whiten") is None else bool(c["whiten"]))
    if name == "MixtureOfSVGPExperts":
        return MixtureOfSVGPExperts(experts_list=[_deserialize(e, rng) for e in c["experts_list"]],
                                    gating_network=_deserialize(c["gating_network"], rng),
                                    num_samples=c.get("num_samples", 1), num_data=c.get("num_data"),
                                    bound=c.get("bound", "further_gating"))
    raise ValueError(f"Unknown layer: {name}")


def _serialize(obj):
    if obj is None:
        return None
    if isinstance(obj, gl.SquaredExponential):
        return {"class_name": "SquaredExponential", "config": {"lengthscales": obj.lengthscales.numpy(), "variance": obj.variance.numpy()}}
    if isinstance(obj, gl.SeparateIndependent):
        return {"class_name": "SeparateIndependent", "config": {"kernels": [_serialize(k) for k in obj.kernels]}}
    if isinstance(obj, gl.SharedIndependent):
        return {"class_name": "SharedIndependent", "config": {"kernel": _serialize(obj.kernel), "output_dim": obj.output_dim}}
    if isinstance(obj, gl.Gaussian):
        return {"class_name": "Gaussian", "config": {"variance": obj.variance.numpy(), "variance_lower_bound": obj.variance_lower_bound}}
    if isinstance(obj, gl.Constant):
        return {"class_name": "Constant", "config": {"c": obj.c.numpy()}}
    if isinstance(obj, gl.Zero):
        return {"class_name": "Zero", "config": {"output_dim": obj.output_dim}}
    if isinstance(obj, gl.InducingPoints):
        return {"class_name": "InducingPoints", "config": {"Z": obj.Z.numpy()}}
    if isinstance(obj, gl.SharedIndependentInducingVariables):
        return {"class_name": "SharedIndependentInducingVariables", "config": {"inducing_variable": _serialize(obj.inducing_variable)}}
    if isinstance(obj, (SVGPExpert, SVGPGatingNetwork)):
        gp = obj.gp
        c = {"kernel": _serialize(gp.kernel), "inducing_variable": _serialize(gp.inducing_variable),
             "mean_function": _serialize(gp.mean_function), "q_diag": gp.q_diag, "q_mu": gp.q_mu.numpy(), "q_sqrt": gp.q_sqrt.numpy(),
             "whiten": gp.whiten}
        if isinstance(obj, SVGPExpert):
            c.update({"likelihood": _serialize(gp.likelihood), "num_latent_gps": gp.num_latent_gps})
        return {"class_name": type(obj).__name__, "config": c}
    if isinstance(obj, MixtureOfSVGPExperts):
        return {"class_name": "MixtureOfSVGPExperts",
                "config": {"experts_list": [_serialize(e) for e in obj.experts_list], "gating_network": _serialize(obj.gating_network),
                           "num_samples": obj.num_samples, "num_data": obj.num_data, "bound": obj.bound}}
    raise TypeError(f"cannot serialise {type(obj).__name__}")


def serialize(obj) -> dict:
    """tf.keras.layers.serialize equivalent for the classes of this package."""
    return _serialize(obj)


def deserialize(cfg: dict, seed=None):
    """tf.keras.layers.deserialize equivalent."""
    return _deserialize(cfg, np.random.default_rng(seed))


def model_from_yaml(yaml_cfg_filename: str, custom_objects: dict = None, seed=None):
    """mogpe/keras/utils.py:36-40."""
    with open(yaml_cfg_filename, "r") as fp:
        cfg = yaml.load(fp, Loader=yaml.SafeLoader)
    return deserialize(cfg, seed)


def save_json_config(obj, filename: str = "config.json"):
    with open(filename, "w") as f:
        json.dump(serialize(obj), f, cls=NumpyArrayEncoder)


def load_from_json_config(filename: str, custom_objects: dict = None):
    with open(filename, "r") as f:
        return deserialize(json.load(f))


def sample_inducing_inputs_from_data(X, num_inducing: int, active_dims: Optional[List[int]] = None, rng=None):
    """Random rows of X (mogpe/keras/utils.py:83-93)."""
    X = np.asarray(X, dtype=np.float64)
    rng = rng or np.random.default_rng()
    idx = rng.choice(X.shape[0], size=num_inducing, replace=False)
    if active_dims is not None:
        X = X[..., active_dims]
    return X[idx, ...].reshape(-1, X.shape[1])


def sample_inducing_variables_from_data(X, inducing_variable, active_dims=None, rng=None):
    Z = inducing_variable.Z
    Z.assign(sample_inducing_inputs_from_data(X, Z.numpy().shape[0], active_dims, rng))


def sample_mosvgpe_inducing_inputs_from_data(X, model: MixtureOfSVGPExperts, seed=None):
    """Re-initialise every inducing input from data rows (mogpe/keras/utils.py:43-52)."""
    rng = np.random.default_rng(seed)
    for expert in model.experts_list:
        sample_inducing_variables_from_data(X, expert.gp.inducing_variable, rng=rng)
    sample_inducing_variables_from_data(X, model.gating_network.gp.inducing_variable, rng=rng)
