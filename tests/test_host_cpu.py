"""CPU-only tests of the host side: bijectors / Parameter bookkeeping, parameter collection with aliasing,
the C-ABI library loads and exports every symbol the header declares, loud failure without a GPU."""
import os
import re

import numpy as np
import pytest

from mogpe_b200 import _lib
from mogpe_b200 import gpflow_lite as gl
from oracle import ref_numpy as rn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    header = open(os.path.join(ROOT, "include", "mogpe_b200.h")).read()
    declared = set(re.findall(r"\b(mogpe_[a-z0-9_]+)\s*\(", header))
    declared -= {"mogpe_status"}
    lib = _lib.load()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert b"sm_100a" in lib.mogpe_version()


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tests._engine_helpers import engine_from_spec
    from tests._specs import make_spec

    spec, X, Y, eps = make_spec()
    with pytest.raises(_lib.MogpeError, match="no CPU fallback"):
        engine_from_spec(spec, 32)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mogpe_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle's", "").replace("the oracle", ""), os.path.join(dirpath, f)


def test_bijectors_match_oracle_definitions():
    rng = np.random.default_rng(0)
    u = rng.standard_normal(7) * 3
    pos = gl.positive(lower=1e-6)
    np.testing.assert_allclose(pos.forward(u), 1e-6 + rn.softplus(u), rtol=1e-15)
    np.testing.assert_allclose(pos.inverse(pos.forward(u)), u, rtol=1e-10, atol=1e-12)
    x = rng.standard_normal((2, 15))
    tri = gl.triangular()
    np.testing.assert_array_equal(tri.forward(x), rn.fill_triangular(x))
    np.testing.assert_array_equal(tri.inverse(tri.forward(x)), x)
    np.testing.assert_array_equal(tri.forward(np.arange(1.0, 7.0)), [[4, 0, 0], [6, 5, 0], [3, 2, 1]])  # TFP docs example
    # chain rule: finite differences through the bijector
    g = rng.standard_normal(7)
    h = 1e-6
    fd = (pos.forward(u + h) - pos.forward(u - h)) / (2 * h) * g
    np.testing.assert_allclose(pos.chain_grad(u, g), fd, rtol=1e-6)
    G = np.tril(rng.standard_normal((2, 5, 5)))
    gu = tri.chain_grad(x, G)
    np.testing.assert_allclose(np.sum(gu * x), np.sum(G * tri.forward(x)), rtol=1e-12)  # linear map: <J^T G, x> = <G, J x>


def test_parameter_roundtrip_and_errors():
    p = gl.Parameter([0.5, 2.0], transform=gl.positive())
    np.testing.assert_allclose(p.numpy(), [0.5, 2.0], rtol=1e-14)
    p.assign([1.0, 3.0])
    np.testing.assert_allclose(p.numpy(), [1.0, 3.0], rtol=1e-14)
    with pytest.raises(ValueError):
        gl.Parameter(-1.0, transform=gl.positive())
    with pytest.raises(ValueError):
        gl.Gaussian(variance=1e-7)  # below the 1e-6 lower bound (mogpe/gps/likelihoods.py:21)
    qd = gl.SVGP(gl.RBF(), gl.Gaussian(), np.zeros((3, 1)), q_diag=True, num_latent_gps=2)  # gpflow: q_sqrt [M, L], positive()
    assert qd.q_sqrt.numpy().shape == (3, 2) and isinstance(qd.q_sqrt.transform, type(gl.positive()))
    np.testing.assert_array_equal(qd.q_sqrt_full(), np.tile(np.eye(3)[None], (2, 1, 1)))
    with pytest.raises(ValueError):
        gl.SVGP(gl.RBF(), gl.Gaussian(), np.zeros((3, 1)), q_diag=True, q_sqrt=np.ones((1, 3, 3)))
    with pytest.raises(NotImplementedError):
        gl.SVGP(gl.RBF(), gl.Gaussian(), np.zeros((3, 1)), whiten=False)
    with pytest.raises(ValueError):
        gl.SVGP(gl.RBF(), gl.Gaussian(), np.zeros((3, 1)), q_mu=np.zeros((4, 1)))


def _reference_fixture_model():
    """tests/keras/test_keras.py:16-108 of the reference: shared kernel / mean / inducing objects, the same
    expert placed in the list twice."""
    from mogpe_b200.keras import MixtureOfSVGPExperts, SVGPExpert, SVGPGatingNetwork

    kern = gl.RBF(lengthscales=2 * [3.0], variance=10.0)
    lik = gl.Gaussian(variance=3.0, variance_lower_bound=1e-6)
    mean = gl.Constant(c=-2.0)
    iv = gl.InducingPoints(np.ones((4, 2)))
    expert = SVGPExpert(kernel=kern, likelihood=lik, mean_function=mean, inducing_variable=iv, num_latent_gps=1,
                        q_diag=False, q_mu=None, q_sqrt=None, whiten=True)
    gating = SVGPGatingNetwork(kernel=kern, mean_function=mean, inducing_variable=iv, q_diag=False, q_mu=None,
                               q_sqrt=None, whiten=True)
    return MixtureOfSVGPExperts(experts_list=[expert for _ in range(2)], gating_network=gating)


def test_reference_fixture_constructs_and_dedupes_aliased_parameters():
    model = _reference_fixture_model()
    assert model.num_experts == 2 and model.gating_network.num_experts == 2
    assert model.bound == "further_gating" and model.num_samples == 1 and model.num_data is None
    names = sorted(p.name for p in model.trainable_variables)
    # kernel (2) + likelihood (1) + Z (1) + c (1) shared; q_mu/q_sqrt of the one expert and of the gating net
    assert names == sorted(["variance", "lengthscales", "variance", "Z", "c", "q_mu", "q_sqrt", "q_mu", "q_sqrt"])
    assert np.array_equal(model.experts_list[0].gp.q_sqrt.numpy()[0], np.eye(4))
    with pytest.raises(AssertionError):
        from mogpe_b200.keras import MixtureOfSVGPExperts

        MixtureOfSVGPExperts(model.experts_list + model.experts_list[:1], model.gating_network)  # 3 experts, Bernoulli gate


def test_gating_likelihood_is_inferred_from_kernel_type():
    from mogpe_b200.keras import SVGPGatingNetwork

    iv = gl.InducingPoints(np.zeros((3, 1)))
    assert isinstance(SVGPGatingNetwork(gl.RBF(), iv).gp.likelihood, gl.Bernoulli)
    g3 = SVGPGatingNetwork(gl.SeparateIndependent([gl.RBF() for _ in range(3)]), iv)
    assert isinstance(g3.gp.likelihood, gl.Softmax) and g3.num_experts == 3 and g3.num_gating_gps == 3
    g4 = SVGPGatingNetwork(gl.SharedIndependent(gl.RBF(), 4), iv)
    assert g4.num_experts == 4 and len(g4.gp.latent_kernel_list) == 1


def test_legacy_constructors_and_validation():
    from mogpe_b200.experts import SVGPExpert, SVGPExperts
    from mogpe_b200.gating_networks import SVGPGatingNetwork
    from mogpe_b200.mixture_of_experts import MixtureOfSVGPExperts

    Z = np.linspace(-1, 1, 6)[:, None]
    experts = SVGPExperts([SVGPExpert(gl.RBF(), gl.Gaussian(0.1), Z.copy()) for _ in range(2)])
    with pytest.raises(AttributeError):
        SVGPGatingNetwork(gl.RBF(), gl.Gaussian(), Z.copy(), None)
    gate = SVGPGatingNetwork(gl.RBF(), gl.Bernoulli(), Z.copy(), None)
    model = MixtureOfSVGPExperts(gate, experts, num_data=100, num_samples=2, bound="further")
    assert model.num_experts == 2 and len(experts.noise_variances()) == 2
    assert len(model.trainable_variables) == 2 * 6 + 5  # per expert: kernel(2) lik Z q_mu q_sqrt; gate: kernel(2) Z q_mu q_sqrt
    u = experts.experts_list[0].sample_inducing_points(3, eps_u=np.zeros((3, 1, 6)))
    assert u.shape == (3, 6, 1) and not np.any(u)
    model.bound = "tight_2"
    with pytest.raises(NotImplementedError):
        model._bound_name()
    with pytest.raises(RuntimeError):
        SVGPExpert(gl.RBF(), gl.Gaussian(0.1), Z.copy()).predict_f(Z)  # not attached to a model


def test_adam_matches_tf_defaults_on_a_quadratic():
    from mogpe_b200._model_core import Adam

    p = gl.Parameter(np.array([1.0, -2.0]))
    opt = Adam(learning_rate=0.1)
    m = v = np.zeros(2)
    x = np.array([1.0, -2.0])
    for t in range(1, 6):
        g = 2 * x
        opt.apply_gradients([(2 * p.unconstrained, p)])
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        x = x - 0.1 * np.sqrt(1 - 0.999**t) / (1 - 0.9**t) * m / (np.sqrt(v) + 1e-7)
        np.testing.assert_allclose(p.unconstrained, x, rtol=1e-12)


def test_mixture_distribution_matches_tfp_formulas():
    from mogpe_b200._model_core import MixtureDistribution

    rng = np.random.default_rng(1)
    pi = rng.dirichlet(np.ones(3), 5)
    mu, sd = rng.standard_normal((5, 2, 3)), rng.uniform(0.5, 2.0, (5, 2, 3))
    mean = (pi[:, None] * mu).sum(-1)
    var = (pi[:, None] * sd**2).sum(-1) + (pi[:, None] * mu**2).sum(-1) - mean**2
    d = MixtureDistribution(pi, mu, sd, mean, var)
    y = rng.standard_normal((5, 2))
    joint = np.log(sum(pi[:, k] * np.prod(np.exp(-0.5 * ((y - mu[..., k]) / sd[..., k]) ** 2) / (sd[..., k] * np.sqrt(2 * np.pi)), -1)
                       for k in range(3)))
    np.testing.assert_allclose(d.log_prob(y), joint, rtol=1e-12)
    assert d.sample(4, rng).shape == (4, 5, 2)


def test_parameter_versions_track_every_change_and_arrays_are_read_only():
    """fit() skips re-packing / re-uploading parameters whose version stamp did not change (MixtureCore._stamp): every
    mutation must therefore go through the setter, and in-place writes must fail loudly instead of being missed."""
    from mogpe_b200.gpflow_lite import Parameter, positive

    p = Parameter(np.array([1.0, 2.0]), transform=positive())
    v0 = p.version
    p.assign(np.array([3.0, 4.0]))
    v1 = p.version
    p.assign_unconstrained(p.unconstrained + 1.0)
    v2 = p.version
    p.unconstrained = np.zeros(2)
    assert v0 < v1 < v2 < p.version
    with pytest.raises(ValueError):
        p.unconstrained[0] = 5.0  # read-only view: cannot bypass the version counter
    q = Parameter(1.0)
    assert q.version > p.version  # one global clock: stamps of different parameters never collide


def test_ctypes_prototypes_have_the_header_arity():
    """Every entry point is bound with as many ctypes argument types as the header declares parameters (ABI drift between
    include/mogpe_b200.h and mogpe_b200/_lib.py would otherwise only show up as corrupted arguments on the GPU box)."""
    header = open(os.path.join(ROOT, "include", "mogpe_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    decls = dict(re.findall(r"\b(mogpe_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S))
    assert len(decls) >= 40
    for name, args in decls.items():
        args = " ".join(args.split())
        n = 0 if args in ("", "void") else args.count(",") + 1
        restype, argtypes = _lib._PROTOS[name]
        assert len(argtypes) == n, (name, n, len(argtypes))


def test_gating_likelihood_maps_closed_forms():
    """Host-side link functions behind predict_mixing_probs_given_h (gpflow Bernoulli / Softmax)."""
    from math import erf, sqrt

    from mogpe_b200 import gpflow_lite as gl

    b = gl.Bernoulli()
    assert b.conditional_mean(0.0) == pytest.approx(0.5, abs=1e-15)
    h = np.array([-30.0, -1.3, 0.2, 30.0])
    want = np.array([0.5 * (1 + erf(x / sqrt(2))) * (1 - 2e-3) + 1e-3 for x in h])
    np.testing.assert_allclose(b.conditional_mean(h), want, rtol=1e-14)
    assert b.conditional_mean(h).min() >= 1e-3 and b.conditional_mean(h).max() <= 1 - 1e-3
    np.testing.assert_allclose(b.conditional_mean(h) + b.conditional_mean(-h), 1.0, rtol=1e-14)  # [h, -h] pairs sum to one
    mu, var = np.array([0.7, -0.4]), np.array([2.0, 0.0])
    p, v = b.predict_mean_and_var(mu, var)
    np.testing.assert_allclose(p, b.conditional_mean(mu / np.sqrt(1 + var)), rtol=1e-15)
    np.testing.assert_allclose(v, p - p * p, rtol=1e-15)
    s = gl.Softmax(3)
    F = np.random.default_rng(0).standard_normal((5, 3)) * 4
    pm = s.conditional_mean(F)
    np.testing.assert_allclose(pm.sum(-1), 1.0, rtol=1e-15)
    np.testing.assert_allclose(pm, np.exp(F) / np.exp(F).sum(-1, keepdims=True), rtol=1e-13)
    eps = np.random.default_rng(1).standard_normal((7, 5, 3))
    m, v = s.predict_mean_and_var(F, np.full((5, 3), 0.3), eps=eps)
    ref = np.stack([s.conditional_mean(F + np.sqrt(0.3) * e) for e in eps])
    np.testing.assert_allclose(m, ref.mean(0), rtol=1e-14)
    np.testing.assert_allclose(v, ref.var(0), rtol=1e-10, atol=1e-15)
    np.testing.assert_allclose(s.predict_mean_and_var(F, np.zeros((5, 3)), eps=eps)[0], pm, rtol=1e-14)  # zero variance


def test_dlpack_capsule_is_consumed_per_protocol_on_the_host():
    """device.as_view on a producer that exposes only __dlpack__ (here: numpy's CPU capsule): zero-copy pointer, capsule
    renamed to "used_dltensor", producer's deleter called once when the view dies (numpy then drops its reference)."""
    import ctypes as C
    import gc
    import sys

    from mogpe_b200 import device as dev

    class OnlyDLPack:
        def __init__(self, a):
            self.a, self.capsules = a, []

        def __dlpack__(self, stream=None):
            self.capsules.append(self.a.__dlpack__())
            return self.capsules[-1]

    a = np.arange(6, dtype=np.float64).reshape(2, 3)
    base_refs = sys.getrefcount(a)
    w = OnlyDLPack(a)
    v = dev.as_view(w)
    assert not v.on_device and v.ptr == a.ctypes.data and v.shape == (2, 3)
    C.pythonapi.PyCapsule_GetName.restype = C.c_char_p
    C.pythonapi.PyCapsule_GetName.argtypes = [C.py_object]
    assert C.pythonapi.PyCapsule_GetName(w.capsules[0]) == b"used_dltensor"
    assert sys.getrefcount(a) > base_refs  # the managed tensor holds the array
    del v
    gc.collect()
    assert sys.getrefcount(a) <= base_refs + 1  # ... and released it through the deleter (w.a is the +1)
    with pytest.raises(TypeError, match="float64"):
        dev.as_view(OnlyDLPack(np.zeros(3, dtype=np.float32)))
    with pytest.raises(ValueError, match="contiguous"):
        dev.as_view(OnlyDLPack(np.zeros((4, 4))[:, ::2]))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the arm the driver runs beside ours): one JSON line with the contract's keys, no GPU needed --
    it times the restated reference (oracle/ref_torch.py) on the host cores."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "C1", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "elbo_fwd_bwd_points_per_sec" and line["unit"] == "points/s"
    assert line["steps"] == 2 and line["warmup"] == 3 and line["value"] > 0 and "workload" in line["config"]  # (W >= 3 is enforced)
    assert line["e2e"] == {"value": line["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
