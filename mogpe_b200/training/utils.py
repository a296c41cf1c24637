"""Checkpoint interop: load the variables of a reference tf.train.Checkpoint(model=model) -- written by the
reference's CheckpointManager, mogpe/training/utils.py:22-43,228-242 -- into a legacy-flavour model built here, so models
trained with the reference can be evaluated (and trained further) on the CUDA path.  No TensorFlow needed."""
import os
import re

import numpy as np

from ..gpflow_lite import Constant, SeparateIndependent
from .tf_checkpoint import read_tf_checkpoint

_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


def _strip(name):
    name = name[:-len(_SUFFIX)] if name.endswith(_SUFFIX) else name
    return name.replace("/_pretransformed_input", "")


def _assign(param, value, what):
    value = np.asarray(value, dtype=np.float64)
    if value.size != param.unconstrained.size:
        raise ValueError(f"{what}: checkpoint has shape {value.shape}, model expects {param.unconstrained.shape}")
    param.assign_unconstrained(value.reshape(param.unconstrained.shape))  # checkpoints store the unconstrained variables


def _load_svgp(gp, prefix, tensors, used):
    def take(key):
        full = f"{prefix}/{key}"
        if full not in tensors:
            return None
        used.add(full)
        return tensors[full]

    z = take("inducing_variable/inducing_variable/Z")
    z = take("inducing_variable/Z") if z is None else z
    if z is not None:
        _assign(gp.Z, z, f"{prefix}/Z")
    kernels = gp.kernel.kernels if isinstance(gp.kernel, SeparateIndependent) else gp.latent_kernel_list
    for i, k in enumerate(kernels):
        for field in ("lengthscales", "variance"):
            v = take(f"kernel/kernels/{i}/{field}")
            if v is None and len(kernels) == 1:
                v = take(f"kernel/{field}")
            if v is None and field == "lengthscales":
                continue
            if v is not None:
                p = getattr(k, field)
                if field == "lengthscales" and p.unconstrained.size == 1 and v.size > 1:
                    raise ValueError(f"{prefix}: checkpoint lengthscales are ARD {v.shape}; build the model with ARD kernels")
                _assign(p, v, f"{prefix}/kernel/{i}/{field}")
    for key, p in (("q_mu", gp.q_mu), ("q_sqrt", gp.q_sqrt)):
        v = take(key)
        if v is not None:
            _assign(p, v, f"{prefix}/{key}")
    v = take("likelihood/variance")
    if v is not None:
        _assign(gp.likelihood.variance, v, f"{prefix}/likelihood/variance")
    v = take("mean_function/c")
    if v is not None and isinstance(gp.mean_function, Constant):
        _assign(gp.mean_function.c, v, f"{prefix}/mean_function/c")


def assign_checkpoint_tensors(model, tensors, strict=True):
    """Assign {checkpoint variable name: array} (names as in the reference's checkpoints, with or without the
    '/.ATTRIBUTES/VARIABLE_VALUE' suffix) to the legacy-flavour `model`."""
    tensors = {_strip(k): v for k, v in tensors.items()}
    used = set()
    for k, expert in enumerate(model.experts.experts_list):
        _load_svgp(expert, f"model/experts/experts_list/{k}", tensors, used)
    _load_svgp(model.gating_network, "model/gating_network", tensors, used)
    unused = [k for k in tensors if k.startswith("model/") and k not in used]
    if strict and unused:
        raise ValueError(f"checkpoint variables without a counterpart in the model: {unused}")
    return sorted(used)


def load_tf_checkpoint_into_model(model, ckpt_prefix, strict=True):
    """Assign every variable of `<ckpt_prefix>.index/.data-*` to the legacy-flavour `model` (same structure as the
    model that was saved).  Returns the list of checkpoint variables that were used."""
    return assign_checkpoint_tensors(model, read_tf_checkpoint(ckpt_prefix), strict)


def update_model_from_checkpoint(model, ckpt_dir):
    """mogpe/training/utils.py:22-28: restore the latest checkpoint of `ckpt_dir` into `model`."""
    state = open(os.path.join(ckpt_dir, "checkpoint")).read()
    m = re.search(r'^model_checkpoint_path:\s*"([^"]+)"', state, re.M)
    if not m:
        raise FileNotFoundError(f"no model_checkpoint_path in {ckpt_dir}/checkpoint")
    load_tf_checkpoint_into_model(model, os.path.join(ckpt_dir, os.path.basename(m.group(1))))
    return model
