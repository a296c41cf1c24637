"""Pins the oracle: two independent implementations agree, analytic known answers hold,
autograd gradients match central finite differences.  CPU only."""
import copy
import math

import numpy as np
import pytest

from oracle import ref_numpy as rn
from oracle import ref_torch as rt
from tests._specs import BOUNDS, FLAVOURS, make_spec


def _tight_ok(spec):
    # Softmax predict_mean_and_var is Monte-Carlo (needs eps_mc); covered separately
    return not (spec["bound"] == "tight" and len(spec["experts"]) > 2)


@pytest.mark.parametrize("flavour", FLAVOURS)
@pytest.mark.parametrize("bound", BOUNDS)
@pytest.mark.parametrize("K,F,S,sep", [(2, 1, 1, False), (2, 2, 2, True), (3, 1, 2, False), (3, 2, 1, True)])
def test_numpy_vs_torch_elbo(flavour, bound, K, F, S, sep):
    spec, X, Y, eps = make_spec(seed=K * 10 + F, N=20, D=2, F=F, K=K, M=7, S=S, flavour=flavour,
                                bound=bound, num_data=137, separate_kernels=sep)
    if not _tight_ok(spec):
        rng = np.random.default_rng(5)
        eps["mc"] = rng.standard_normal((S, 11, 20, K))
    a = rn.elbo(spec, X, Y, eps)
    b = float(rt.MixtureOfSVGPExperts(spec).elbo(X, Y, eps).detach())
    assert np.isfinite(a)
    assert abs(a - b) <= 1e-11 * max(1.0, abs(a)), (a, b)


@pytest.mark.parametrize("flavour", FLAVOURS)
@pytest.mark.parametrize("K", [2, 3])
def test_numpy_vs_torch_predict(flavour, K):
    spec, X, Y, eps = make_spec(seed=3, N=15, D=3, F=2, K=K, M=6, flavour=flavour, separate_kernels=True)
    eps_mc = np.random.default_rng(1).standard_normal((13, 15, K)) if K > 2 else None
    m = rt.MixtureOfSVGPExperts(spec)
    pa = rn.predict_mixing_probs(spec, X, eps_mc)
    pb = m.predict_mixing_probs(X, eps_mc).detach().numpy()
    np.testing.assert_allclose(pa, pb, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(pa.sum(-1), 1.0, rtol=0, atol=1e-12)
    ma, va = rn.predict_y(spec, X, eps_mc)
    mb, vb = m.predict_y(X, eps_mc)
    np.testing.assert_allclose(ma, mb.detach().numpy(), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(va, vb.detach().numpy(), rtol=1e-10, atol=1e-13)


def test_fill_triangular_tfp_docs_example():
    # tfp.math.fill_triangular([1,2,3,4,5,6]) == [[4,0,0],[6,5,0],[3,2,1]]
    L = rn.fill_triangular(np.arange(1.0, 7.0))
    np.testing.assert_array_equal(L, [[4, 0, 0], [6, 5, 0], [3, 2, 1]])
    np.testing.assert_array_equal(rn.fill_triangular_inverse(L), np.arange(1.0, 7.0))
    Lt = rt.fill_triangular(rt._t(np.arange(1.0, 7.0))).numpy()
    np.testing.assert_array_equal(Lt, L)
    x = np.random.default_rng(0).standard_normal((3, 15))
    np.testing.assert_array_equal(rn.fill_triangular_inverse(rn.fill_triangular(x)), x)


def test_known_answer_default_init_kl_zero_and_mean_is_c():
    spec, X, Y, eps = make_spec(seed=1, N=10, D=2, F=1, K=2, M=5, mean_c=-2.0)
    for gp in spec["experts"] + [spec["gating"]]:
        gp["q_mu"] = np.zeros_like(gp["q_mu"])
        gp["q_sqrt"] = np.broadcast_to(np.eye(5), gp["q_sqrt"].shape).copy()
        assert rn.prior_kl(gp) == 0.0
    f_mean, _ = rn.predict_f(spec["experts"][0], X)
    np.testing.assert_allclose(f_mean, np.asarray(spec["experts"][0]["mean_c"]).reshape(1, -1).repeat(10, 0), rtol=0, atol=0)
    # gating: h_mean = 0 (zero mean function) -> pi = [1/2, 1/2]
    pi = rn.predict_mixing_probs(spec, X)
    np.testing.assert_allclose(pi, 0.5, rtol=0, atol=1e-15)


def test_known_answer_reference_fixture_geometry():
    """tests/keras/test_keras.py:16-49 geometry: X = Z = 1, RBF(ls=[3,3], var=10), Gaussian(3), Constant(-2).
    Kuu = 10*11^T + 1e-6 I; closed forms: with a = Kuf column = 10*1, A^T A = a^T Kuu^-1 a = 40*10/(40+1e-6)...
    """
    N, M, D = 5, 4, 2
    kern = {"lengthscales": [3.0, 3.0], "variance": 10.0}
    gp = {"Z": np.ones((M, D)), "kernels": [kern], "q_mu": np.zeros((M, 1)),
          "q_sqrt": np.eye(M)[None].copy(), "mean_c": [-2.0], "noise_variance": [3.0]}
    gate = copy.deepcopy(gp)
    del gate["noise_variance"]
    X, Y = np.ones((N, D)), np.ones((N, 1))
    Kuu = rn.rbf_K(kern, gp["Z"]) + 1e-6 * np.eye(M)
    np.testing.assert_allclose(Kuu, 10.0 * np.ones((M, M)) + 1e-6 * np.eye(M), rtol=0, atol=1e-14)
    mean, var = rn.predict_f(gp, X)
    # Kuu^-1 1 = 1 / (10 M + 1e-6)  =>  sum A^2 = 100 M / (10 M + 1e-6); q_sqrt = I adds the same back
    s = 100.0 * M / (10.0 * M + 1e-6)
    np.testing.assert_allclose(var, 10.0 - s + s, rtol=1e-9)
    np.testing.assert_allclose(mean, -2.0)
    # same expert object twice (reference fixture aliases it) -> K = 2 identical experts
    spec = {"flavour": "keras", "bound": "further_gating", "num_samples": 1, "num_data": None,
            "experts": [gp, copy.deepcopy(gp)], "gating": gate}
    eps = {"u_experts": [np.zeros((1, 1, M))] * 2, "h": np.zeros((1, N, 1))}
    val = rn.elbo(spec, X, Y, eps)
    # u = 0 -> f_mean = -2, f_var = 10 - s, y_var = f_var + 3 used as *scale* (keras quirk)
    sc = 10.0 - s + 3.0
    p = math.exp(-0.5 * (3.0 / sc) ** 2) / (sc * math.sqrt(2 * math.pi))
    pi = rn.inv_probit(np.array(-2.0))  # h = mean_c = -2, eps = 0
    expect = N * math.log(p * pi + p * (1 - pi))
    assert abs(val - expect) < 1e-9


def test_inducing_variance_is_sample_invariant_and_has_no_q_sqrt_term():
    spec, X, Y, eps = make_spec(seed=2, N=9, D=1, F=1, K=2, M=6, S=3)
    gp = spec["experts"][0]
    mean, var = rn.predict_f_given_inducing_samples(gp, X, eps["u_experts"][0])
    assert np.all(var[0] == var[1]) and np.all(var[1] == var[2])
    _, var_marg = rn.predict_f(gp, X)
    assert np.all(var_marg > var[0])  # the marginal adds sum (q_sqrt^T A)^2 > 0


def test_bernoulli_probs_range_and_sum():
    spec, X, Y, eps = make_spec(seed=4, N=50, D=1, K=2, M=6)
    for fl in FLAVOURS:
        spec["flavour"] = fl
        h = 50.0 * np.random.default_rng(0).standard_normal((1, 50, 1))
        p = rn._mixing_probs_given_h(spec, h)
        assert p.min() >= 1e-3 - 1e-16 and p.max() <= 1 - 1e-3 + 1e-16
        np.testing.assert_allclose(p.sum(-1), 1.0, rtol=0, atol=2e-16)


def test_num_data_scaling():
    spec, X, Y, eps = make_spec(seed=6, N=12, K=2, M=5, noise_lo=0.3)
    v1 = rn.elbo(spec, X, Y, eps)
    assert np.isfinite(v1)
    spec["num_data"] = 12
    assert rn.elbo(spec, X, Y, eps) == pytest.approx(v1, rel=1e-15)
    spec["num_data"] = 120
    kl = sum(rn.prior_kl(g) for g in spec["experts"]) + rn.prior_kl(spec["gating"])
    assert rn.elbo(spec, X, Y, eps) == pytest.approx((v1 + kl) * 10 - kl, rel=1e-12)


def _perturb(spec, name, idx, h):
    """Return a copy of spec with one *unconstrained* coordinate moved by h."""
    s = copy.deepcopy(spec)
    parts = name.split("/")
    gp = s["gating"] if parts[0] == "gating" else s["experts"][int(parts[1])]
    key = parts[-1] if parts[0] == "gating" else "/".join(parts[2:])
    key = "/".join(parts[1:]) if parts[0] == "gating" else key
    if key == "Z":
        gp["Z"] = np.array(gp["Z"], dtype=float)
        gp["Z"][idx] += h
    elif key == "q_mu":
        gp["q_mu"][idx] += h
    elif key == "q_sqrt":
        u = rn.fill_triangular_inverse(np.tril(gp["q_sqrt"]))
        u[idx] += h
        gp["q_sqrt"] = rn.fill_triangular(u)
    elif key == "mean_c":
        gp["mean_c"] = np.array(gp["mean_c"], dtype=float)
        gp["mean_c"][idx] += h
    elif key == "noise_variance":
        v = np.array(gp["noise_variance"], dtype=float)
        u = rn.softplus_inverse(v - 1e-6)
        u[idx] += h
        gp["noise_variance"] = 1e-6 + rn.softplus(u)
    elif key.startswith("kernels/"):
        ki = int(key.split("/")[1])
        field = key.split("/")[2]
        kern = gp["kernels"][ki]
        v = np.array(kern[field], dtype=float)
        u = rn.softplus_inverse(v)
        if u.ndim == 0:
            u = u + h
        else:
            u[idx] += h
        kern[field] = rn.softplus(u)
    else:
        raise KeyError(name)
    return s


@pytest.mark.parametrize("flavour", FLAVOURS)
@pytest.mark.parametrize("bound", BOUNDS)
@pytest.mark.parametrize("K", [2, 3])
def test_autograd_vs_finite_differences(flavour, bound, K):
    spec, X, Y, eps = make_spec(seed=11 + K, N=10, D=2, F=1, K=K, M=5, S=2, flavour=flavour, bound=bound,
                                num_data=40, gating_mean_c=0.1)
    if not _tight_ok(spec):
        eps["mc"] = np.random.default_rng(5).standard_normal((2, 7, 10, K))
    model = rt.MixtureOfSVGPExperts(spec)
    val, grads = model.elbo_and_grads(X, Y, eps)
    rng = np.random.default_rng(0)
    h = 1e-6
    for name, g in grads.items():
        g = np.asarray(g)
        for _ in range(2):
            idx = tuple(int(rng.integers(0, d)) for d in g.shape)
            fd = (rn.elbo(_perturb(spec, name, idx, h), X, Y, eps) - rn.elbo(_perturb(spec, name, idx, -h), X, Y, eps)) / (2 * h)
            gi = g[idx] if g.shape else float(g)
            assert abs(fd - gi) <= 2e-6 * max(1.0, abs(gi), abs(val) * 1e-3), (name, idx, fd, gi)


def test_active_dims_slice_the_kernel_inputs():
    """gpflow Kernel.slice: a kernel with active_dims is the same function of the selected input columns; both oracle
    implementations honour it and agree, and the inducing-input gradient of an inactive column is exactly zero."""
    from tests._specs import restrict_active_dims

    spec, X, Y, eps = make_spec(seed=12, N=24, D=3, F=1, K=2, M=6, S=2, flavour="keras", num_data=200)
    restrict_active_dims(spec, expert_dims=[0, 2], gating_dims=[1])
    k = spec["experts"][0]["kernels"][0]
    Z = spec["experts"][0]["Z"]
    plain = {"lengthscales": k["lengthscales"], "variance": k["variance"]}
    np.testing.assert_array_equal(rn.rbf_K(k, X, Z), rn.rbf_K(plain, X[:, [0, 2]], Z[:, [0, 2]]))
    a = rn.elbo(spec, X, Y, eps)
    val, grads = rt.MixtureOfSVGPExperts(spec).elbo_and_grads(X, Y, eps)
    assert abs(a - float(val)) <= 1e-11 * max(1.0, abs(a))
    assert np.all(np.asarray(grads["experts/0/Z"])[:, 1] == 0.0) and np.all(np.asarray(grads["gating/Z"])[:, [0, 2]] == 0.0)
    assert np.asarray(grads["experts/0/kernels/0/lengthscales"]).shape == (2,)


def test_front_end_kernels_accept_active_dims():
    """gpflow_lite.RBF(active_dims=...) (list or slice) resolves to column indices; ARD lengthscales must match their count."""
    from mogpe_b200 import gpflow_lite as gl

    assert gl.RBF(lengthscales=[1.0, 2.0], active_dims=[0, 2]).active_columns(3).tolist() == [0, 2]
    assert gl.RBF(active_dims=slice(1, 3)).active_columns(4).tolist() == [1, 2]
    assert gl.RBF(lengthscales=[1.0, 2.0], active_dims=[0, 1]).active_columns(2) is None  # all columns, in order
    with pytest.raises(ValueError):
        gl.RBF(lengthscales=[1.0, 2.0, 3.0], active_dims=[0, 2])
    with pytest.raises(ValueError):
        gl.RBF(active_dims=[0, 5]).active_columns(3)


def test_q_diag_is_the_diagonal_special_case():
    """gpflow q_diag=True (q_sqrt [M, L], positive transform): the same bound as full factors that happen to be diagonal, in
    both oracle implementations; the unconstrained q_sqrt gradient is the softplus-chained diagonal of the full one."""
    from tests._specs import make_q_diag

    spec, X, Y, eps = make_spec(seed=21, N=20, D=2, F=2, K=2, M=6, S=2, flavour="keras", bound="further", num_data=150,
                                separate_kernels=True)
    full = copy.deepcopy(spec)
    make_q_diag(spec)
    for gp in full["experts"] + [full["gating"]]:
        d = np.abs(np.diagonal(np.asarray(gp["q_sqrt"]), axis1=-2, axis2=-1))
        gp["q_sqrt"] = np.stack([np.diag(r) for r in d])
    a, b = rn.elbo(spec, X, Y, eps), rn.elbo(full, X, Y, eps)
    assert a == pytest.approx(b, rel=1e-13)
    val, grads = rt.MixtureOfSVGPExperts(spec).elbo_and_grads(X, Y, eps)
    assert abs(a - float(val)) <= 1e-11 * max(1.0, abs(a))
    assert np.asarray(grads["experts/0/q_sqrt"]).shape == (6, 2) and np.asarray(grads["gating/q_sqrt"]).shape == (6, 1)
