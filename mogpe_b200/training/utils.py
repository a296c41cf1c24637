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


# ---- saving (mogpe/training/utils.py:228-242: tf.train.Checkpoint(model=model) + CheckpointManager) ---------------------
def _svgp_tensors(gp, prefix, out):
    """The inverse of _load_svgp: {reference checkpoint key: unconstrained variable} of one SVGP."""
    from ..gpflow_lite import SharedIndependentInducingVariables

    def put(key, param, shape=None):
        v = np.asarray(param.unconstrained, dtype=np.float64)
        out[f"{prefix}/{key}/_pretransformed_input{_SUFFIX}"] = v.reshape(shape) if shape is not None else v

    iv = getattr(gp, "inducing_variable", None)
    shared = isinstance(iv, SharedIndependentInducingVariables)
    put("inducing_variable/inducing_variable/Z" if shared else "inducing_variable/Z", gp.Z)
    if isinstance(gp.kernel, SeparateIndependent):
        for i, k in enumerate(gp.kernel.kernels):
            put(f"kernel/kernels/{i}/lengthscales", k.lengthscales)
            put(f"kernel/kernels/{i}/variance", k.variance, (1,))  # the reference builds [1]-shaped variances (model_parsers.py:30-35)
    else:
        k = gp.latent_kernel_list[0]
        put("kernel/lengthscales", k.lengthscales)
        put("kernel/variance", k.variance, (1,))
    put("q_mu", gp.q_mu)
    put("q_sqrt", gp.q_sqrt)
    lik = getattr(gp, "likelihood", None)
    if lik is not None and hasattr(lik, "variance"):
        put("likelihood/variance", lik.variance)
    if isinstance(gp.mean_function, Constant):
        put("mean_function/c", gp.mean_function.c)


def model_checkpoint_tensors(model):
    """{checkpoint key: array} of a legacy-flavour model, with the variable names of the reference's checkpoints."""
    out = {}
    for k, expert in enumerate(model.experts.experts_list):
        _svgp_tensors(expert, f"model/experts/experts_list/{k}", out)
    _svgp_tensors(model.gating_network, "model/gating_network", out)
    return out


def save_tf_checkpoint(model, ckpt_prefix, save_counter=1):
    """Write the model's variables as a tf.train.Checkpoint(model=model) file pair (no TensorFlow needed); the reference
    restores it with `tf.train.Checkpoint(model=model).restore(ckpt_prefix)`."""
    from .tf_checkpoint import write_tf_checkpoint

    return write_tf_checkpoint(ckpt_prefix, model_checkpoint_tensors(model), save_counter=save_counter)


class CheckpointManager:
    """tf.train.CheckpointManager(ckpt, log_dir, max_to_keep): numbered `ckpt-N` files plus the `checkpoint` state file."""

    def __init__(self, model, directory, max_to_keep=5):
        self.model, self.directory, self.max_to_keep = model, directory, max_to_keep
        os.makedirs(directory, exist_ok=True)
        self.checkpoints, self.save_counter = [], 0

    @property
    def latest_checkpoint(self):
        return self.checkpoints[-1] if self.checkpoints else None

    def save(self, checkpoint_number=None):
        self.save_counter += 1
        n = self.save_counter if checkpoint_number is None else int(checkpoint_number)
        prefix = os.path.join(self.directory, f"ckpt-{n}")
        save_tf_checkpoint(self.model, prefix, save_counter=self.save_counter)
        self.checkpoints.append(prefix)
        while self.max_to_keep and len(self.checkpoints) > self.max_to_keep:
            old = self.checkpoints.pop(0)
            for ext in (".index", ".data-00000-of-00001"):
                if os.path.exists(old + ext):
                    os.remove(old + ext)
        with open(os.path.join(self.directory, "checkpoint"), "w") as f:
            f.write(f'model_checkpoint_path: "{os.path.basename(prefix)}"\n')
            for c in self.checkpoints:
                f.write(f'all_model_checkpoint_paths: "{os.path.basename(c)}"\n')
        return prefix


def init_checkpoint_manager(model, log_dir, num_ckpts, config_file=None, train_dataset=None, test_dataset=None):
    """mogpe/training/utils.py:228-242 with the same side effects (config copy, data set snapshots)."""
    import shutil

    os.makedirs(log_dir, exist_ok=True)
    if config_file is not None:
        try:
            shutil.copy(config_file, os.path.join(log_dir, "config.toml"))
        except OSError:
            print("Failed to copy config_file to log_dir")
    if train_dataset is not None:
        np.savez(os.path.join(log_dir, "train_dataset"), x=train_dataset[0], y=train_dataset[1])
    if test_dataset is not None:
        np.savez(os.path.join(log_dir, "test_dataset"), x=test_dataset[0], y=test_dataset[1])
    return CheckpointManager(model, log_dir, max_to_keep=num_ckpts)
