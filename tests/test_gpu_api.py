"""GPU parity of the drop-in Python API (both flavours) against the oracle: what a user of the reference calls."""
import numpy as np
import pytest

from tests._specs import BOUNDS, FLAVOURS, make_spec

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


@pytest.mark.parametrize("flavour", FLAVOURS)
@pytest.mark.parametrize("bound", BOUNDS)
@pytest.mark.parametrize("K,F,sep", [(2, 1, False), (2, 2, True), (3, 1, True)])
def test_objective_and_unconstrained_gradients(flavour, bound, K, F, sep):
    from oracle import ref_torch as rt
    from tests._model_helpers import model_from_spec, oracle_grads_in_model_order

    spec, X, Y, eps = make_spec(seed=40 + K + F, N=90, D=2, F=F, K=K, M=14, S=2, flavour=flavour, bound=bound,
                                num_data=900, separate_kernels=sep, gating_mean_c=-0.1)
    if bound == "tight" and K > 2:
        eps["mc"] = np.random.default_rng(3).standard_normal((2, 11, 90, K))
    model = model_from_spec(spec)
    val = model.maximum_log_likelihood_objective((X, Y), eps=eps)
    ref, grads_u = rt.MixtureOfSVGPExperts(spec).elbo_and_grads(X, Y, eps)
    assert abs(val - ref) <= 1e-8 * abs(ref)
    val2, grads = model.elbo_and_gradients((X, Y), eps=eps)
    assert val2 == pytest.approx(val, rel=1e-12)
    want = oracle_grads_in_model_order(spec, model, grads_u)
    for p, g, w in zip(model.trainable_variables, grads, want):
        w = np.asarray(w)
        assert np.max(np.abs(np.asarray(g).reshape(w.shape) - w)) <= 1e-6 * max(1.0, np.max(np.abs(w))), p.name


@pytest.mark.parametrize("flavour", FLAVOURS)
@pytest.mark.parametrize("K", [2, 3])
def test_predict_surface(flavour, K):
    from oracle import ref_numpy as rn
    from tests._model_helpers import model_from_spec

    spec, X, Y, eps = make_spec(seed=60 + K, N=70, D=1, F=2, K=K, M=9, flavour=flavour, separate_kernels=True)
    model = model_from_spec(spec)
    eps_mc = np.random.default_rng(0).standard_normal((23, 70, K)) if K > 2 else None
    dist = model.predict_y(X, eps_mc=eps_mc)
    ym, yv = rn.predict_y(spec, X, eps_mc)
    assert _rel(dist.mean(), ym) < 1e-8 and _rel(dist.variance(), yv) < 1e-8
    pi = rn.predict_mixing_probs(spec, X, eps_mc)
    assert _rel(model.gating_network.predict_mixing_probs(X, eps_mc=eps_mc), pi) < 1e-8
    mu_ref, var_ref = rn.predict_experts_y(spec, X)
    if flavour == "keras":
        mean, var = model(X) if K == 2 else (dist.mean(), dist.variance())
        assert _rel(mean, ym) < 1e-8
        mus, vs = model.predict_experts_y(X)
        e0 = model.experts_list[0]
        d0 = e0.predict_dist(X)
        assert _rel(d0.mean(), mu_ref[..., 0]) < 1e-8 and _rel(d0.stddev(), var_ref[..., 0]) < 1e-8  # quirk 1
        kl = e0.prior_kl()
        gate_kl = model.gating_network.prior_kl()
        h_mean, h_var = model.gating_network.predict_h(X)
        assert h_mean.shape == (70, K) and (K > 2 or np.array_equal(h_mean[:, 1], -h_mean[:, 0]))
    else:
        mus, vs = model.experts.predict_ys(X)
        kl = model.experts.prior_kls()[0]
        gate_kl = model.gating_network.prior_kl()
        assert model.predict_mixing_probs(X, eps_mc=eps_mc).shape == (70, K)
        fm, fv = model.experts.predict_fs(X)
        assert fm.shape == (70, 2, K)
    assert _rel(mus, mu_ref) < 1e-8 and _rel(vs, var_ref) < 1e-8
    assert kl == pytest.approx(rn.prior_kl(spec["experts"][0]), rel=1e-10)
    assert gate_kl == pytest.approx(rn.prior_kl(spec["gating"]), rel=1e-10)
    # log_prob used by the reference's NLPD metric (training/metrics.py:9-16)
    assert np.all(np.isfinite(dist.log_prob(Y, joint_over_outputs=(flavour == "keras"))))


@pytest.mark.parametrize("flavour", FLAVOURS)
def test_predict_f_given_inducing_samples(flavour):
    from oracle import ref_numpy as rn
    from tests._model_helpers import model_from_spec

    spec, X, Y, eps = make_spec(seed=71, N=50, D=2, F=2, K=2, M=11, S=3, flavour=flavour, separate_kernels=True)
    model = model_from_spec(spec)
    expert = model.experts_list[1] if flavour == "keras" else model.experts.experts_list[1]
    mean, var = expert.predict_f(X, num_inducing_samples=3, eps_u=eps["u_experts"][1])
    m_ref, v_ref = rn.predict_f_given_inducing_samples(spec["experts"][1], X, eps["u_experts"][1])
    assert mean.shape == (3, 50, 2) and _rel(mean, m_ref) < 1e-8 and _rel(var, v_ref) < 1e-8


def test_reference_fixture_with_aliased_expert_accumulates_gradients():
    """tests/keras/test_keras.py fixture: same expert twice, shared kernel / mean / Z between expert and gate."""
    import copy

    from oracle import ref_numpy as rn
    from tests.test_host_cpu import _reference_fixture_model

    model = _reference_fixture_model()
    N, M = 5, 4
    X, Y = np.ones((N, 2)), np.ones((N, 1))
    eps = {"u_experts": [np.full((1, 1, M), 0.3)] * 2, "h": np.full((1, N, 1), -0.2)}
    kern = {"lengthscales": [3.0, 3.0], "variance": 10.0}
    gp = {"Z": np.ones((M, 2)), "kernels": [kern], "q_mu": np.zeros((M, 1)), "q_sqrt": np.eye(M)[None].copy(),
          "mean_c": [-2.0], "noise_variance": [3.0]}
    gate = copy.deepcopy(gp)
    del gate["noise_variance"]
    spec = {"flavour": "keras", "bound": "further_gating", "num_samples": 1, "num_data": None,
            "experts": [gp, copy.deepcopy(gp)], "gating": gate}
    ref = rn.elbo(spec, X, Y, eps)
    val, grads = model.elbo_and_gradients((X, Y), eps=eps)
    assert abs(val - ref) <= 1e-8 * abs(ref)
    # finite differences on the shared likelihood variance: both experts' contributions must be summed
    p = model.experts_list[0].gp.likelihood.variance
    idx = [id(v) for v in model.trainable_variables].index(id(p))
    u0 = p.unconstrained.copy()
    h = 1e-5
    p.assign_unconstrained(u0 + h)
    vp = model.maximum_log_likelihood_objective((X, Y), eps=eps)
    p.assign_unconstrained(u0 - h)
    vm = model.maximum_log_likelihood_objective((X, Y), eps=eps)
    p.assign_unconstrained(u0)
    assert float(np.asarray(grads[idx]).reshape(-1)[0]) == pytest.approx((vp - vm) / (2 * h), rel=1e-5)


@pytest.mark.parametrize("flavour", FLAVOURS)
def test_training_reduces_the_loss(flavour):
    from mogpe_b200._model_core import Adam
    from tests._model_helpers import model_from_spec

    spec, X, Y, eps = make_spec(seed=80, N=128, D=1, F=1, K=2, M=10, S=1, flavour=flavour, num_data=128)
    model = model_from_spec(spec)
    model.set_seed(1)
    model.compile(optimizer=Adam(learning_rate=0.01))
    first = model.test_step((X, Y), eps=eps)["loss"] if flavour == "keras" else model.training_loss((X, Y), eps=eps)
    if flavour == "keras":
        hist = model.fit(X, Y, epochs=30, batch_size=32)
        assert len(hist["loss"]) == 30
    else:
        for _ in range(120):
            logs = model.train_step((X, Y))
        assert "loss" in logs
    last = -model.maximum_log_likelihood_objective((X, Y), eps=eps)
    assert np.isfinite(last) and last < first - 1.0, (first, last)


def test_library_random_stream_replays_through_the_oracle():
    """eps left to the library's Philox stream: dump the draws and replay them in the oracle."""
    from oracle import ref_numpy as rn
    from tests._engine_helpers import engine_from_spec

    spec, X, Y, eps = make_spec(seed=90, N=64, D=2, F=1, K=2, M=8, S=2, num_data=640)
    eng = engine_from_spec(spec, max_batch=64)
    eng.set_seed(77)
    res = eng.elbo(X, Y, "further_gating", 2, 10.0, want_grads=False)
    eu = eng.debug("eps_u").reshape(eng.P, 2, eng.Mp)
    eh = eng.debug("eps_h").reshape(2, 64, 1)
    replay = {"u_experts": [np.transpose(eu[k:k + 1, :, :8], (1, 0, 2)) for k in range(2)], "h": eh}
    ref = rn.elbo(spec, X, Y, replay)
    assert abs(res.elbo - ref) <= 1e-8 * abs(ref)
    res2 = eng.elbo(X, Y, "further_gating", 2, 10.0, want_grads=False)
    assert res2.elbo != res.elbo  # the stream advances


@pytest.mark.parametrize("flavour", FLAVOURS)
def test_device_optimizer_matches_host_adam(flavour):
    """fit(device_optimizer=True): bijector chain + Adam on the GPU == train_step's host Adam, step for step
    (same library random stream), including tied variables (scalar lengthscale broadcast over D, shared noise)."""
    from mogpe_b200._model_core import Adam
    from tests._model_helpers import model_from_spec

    spec, X, Y, eps = make_spec(seed=81, N=96, D=2, F=2, K=2, M=9, S=2, flavour=flavour, num_data=96, ard=False,
                                separate_kernels=False, gating_mean_c=0.05)
    a, b = model_from_spec(spec), model_from_spec(spec)
    for m in (a, b):
        m.set_seed(5)
        m.compile(optimizer=Adam(learning_rate=0.02))
    steps = 6
    for _ in range(steps):
        a.train_step((X, Y))                      # host: unpack grads, chain, numpy Adam, re-pack
    b.fit(X, Y, epochs=steps, batch_size=96, shuffle=False)  # device: one batch per epoch = same 6 steps
    assert len(a.trainable_variables) == len(b.trainable_variables)
    for pa, pb in zip(a.trainable_variables, b.trainable_variables):
        np.testing.assert_allclose(pb.unconstrained, pa.unconstrained, rtol=1e-9, atol=1e-11, err_msg=pa.name)
    va = a.maximum_log_likelihood_objective((X, Y), eps=eps)
    vb = b.maximum_log_likelihood_objective((X, Y), eps=eps)
    assert va == pytest.approx(vb, rel=1e-9)


def test_device_resident_data_pipeline_matches_host_batches():
    """fit(device_data=True): shuffled minibatches gathered on the device give the same training trajectory as
    host-side slicing (same permutation stream, same library random stream)."""
    from mogpe_b200._model_core import Adam
    from tests._model_helpers import model_from_spec

    spec, X, Y, eps = make_spec(seed=83, N=100, D=2, F=1, K=2, M=8, S=1, num_data=100)
    a, b = model_from_spec(spec), model_from_spec(spec)
    for m in (a, b):
        m.set_seed(9)
        m.compile(optimizer=Adam(learning_rate=0.01))
    ha = a.fit(X, Y, epochs=3, batch_size=32, shuffle=True, seed=4)                    # 4 batches / epoch, last one ragged
    hb = b.fit(X, Y, epochs=3, batch_size=32, shuffle=True, seed=4, device_data=True)
    np.testing.assert_allclose(hb["loss"], ha["loss"], rtol=1e-10)
    for pa, pb in zip(a.trainable_variables, b.trainable_variables):
        np.testing.assert_allclose(pb.unconstrained, pa.unconstrained, rtol=1e-9, atol=1e-11, err_msg=pa.name)


@pytest.mark.parametrize("device_data", [False, True])
def test_graph_replayed_training_steps_match_direct_launches(device_data):
    """fit() replays each step as one CUDA graph (mogpe_train_step); the trajectory must equal the directly launched one:
    the random-stream position and the Adam step count live on the device and continue across replays, ragged last
    batches get their own graph, and a predict call in the middle (call-mode change) does not disturb the replays."""
    from mogpe_b200._model_core import Adam
    from tests._model_helpers import model_from_spec

    spec, X, Y, eps = make_spec(seed=85, N=100, D=2, F=1, K=2, M=8, S=2, num_data=100)
    a, b = model_from_spec(spec), model_from_spec(spec)
    for m in (a, b):
        m.set_seed(11)
        m.compile(optimizer=Adam(learning_rate=0.01))
    a._core.use_graphs = False
    ha = a.fit(X, Y, epochs=4, batch_size=32, shuffle=False, device_data=device_data)  # 3 full batches + one of 4 rows
    hb = b.fit(X, Y, epochs=2, batch_size=32, shuffle=False, device_data=device_data)
    b.predict_y(X[:10])
    hb2 = b.fit(X, Y, epochs=2, batch_size=32, shuffle=False, device_data=device_data)
    a.predict_y(X[:10])
    np.testing.assert_allclose(hb["loss"] + hb2["loss"], ha["loss"], rtol=1e-10)
    for pa, pb in zip(a.trainable_variables, b.trainable_variables):
        np.testing.assert_allclose(pb.unconstrained, pa.unconstrained, rtol=1e-9, atol=1e-11, err_msg=pa.name)
    captured, disabled = b._core.engine.graph_stats()
    assert not disabled and captured >= 1
    assert a._core.engine.graph_stats()[0] == 0


def test_fit_defers_the_variable_copy_back_until_a_parameter_is_read():
    """fit() leaves the trained variables on the device (no 34 MB D2H per call at the bench shape); the first read of ANY
    Parameter completes the copy, an assignment made while it is pending is not overwritten, and a second fit() in between
    continues from the device state."""
    from mogpe_b200._model_core import Adam
    from mogpe_b200.gpflow_lite import Parameter
    from tests._model_helpers import model_from_spec

    spec, X, Y, eps = make_spec(seed=91, N=64, D=2, F=1, K=2, M=8, S=1, num_data=64)
    a, b = model_from_spec(spec), model_from_spec(spec)
    for m in (a, b):
        m.set_seed(3)
        m.compile(optimizer=Adam(learning_rate=0.02))
    a.fit(X, Y, epochs=4, batch_size=64, shuffle=False)
    b.fit(X, Y, epochs=2, batch_size=64, shuffle=False)
    assert Parameter._pending_pulls  # nothing was copied back yet
    b.fit(X, Y, epochs=2, batch_size=64, shuffle=False)  # continues on the device
    q0 = spec["experts"][0]["q_mu"]
    got = b.experts_list[0].gp.q_mu.numpy() if hasattr(b.experts_list[0], "gp") else b.trainable_variables[0].numpy()
    assert not Parameter._pending_pulls
    for pa, pb in zip(a.trainable_variables, b.trainable_variables):
        np.testing.assert_allclose(pb.unconstrained, pa.unconstrained, rtol=1e-9, atol=1e-11, err_msg=pa.name)
    assert not np.allclose(np.asarray(got).reshape(-1)[:4], np.asarray(q0).reshape(-1)[:4])
    # an assignment while a copy-back is pending wins
    b.fit(X, Y, epochs=1, batch_size=64, shuffle=False)
    p = b.trainable_variables[0]
    new = np.full(p.unconstrained.shape, 0.25)
    p.assign_unconstrained(new)
    np.testing.assert_array_equal(p.unconstrained, new)
    v1 = b.maximum_log_likelihood_objective((X, Y), eps=eps)  # engine re-packs from the Parameters (including `new`)
    p.assign_unconstrained(new + 0.5)
    assert b.maximum_log_likelihood_objective((X, Y), eps=eps) != v1


def test_trainable_flags_changed_between_fits_are_honoured():
    """gpflow.set_trainable staging (the reference's examples freeze / unfreeze parameters between training phases): the
    device-side variable map follows the flags, and Adam's moments of the variables that stay trainable carry over."""
    from mogpe_b200._model_core import Adam
    from tests._model_helpers import model_from_spec

    spec, X, Y, eps = make_spec(seed=93, N=64, D=2, F=1, K=2, M=8, S=1, num_data=64)
    m = model_from_spec(spec)
    m.set_seed(1)
    m.compile(optimizer=Adam(learning_rate=0.05))
    m.fit(X, Y, epochs=2, batch_size=64, shuffle=False)
    params = {name: p for name, p in m._core.gating_gp.parameters()}
    frozen = params["Z"] if "Z" in params else m._core.gating_gp.Z
    before_frozen = frozen.unconstrained.copy()
    others_before = [p.unconstrained.copy() for p in m.trainable_variables if p is not frozen]
    frozen.trainable = False
    m.fit(X, Y, epochs=2, batch_size=64, shuffle=False)
    np.testing.assert_array_equal(frozen.unconstrained, before_frozen)  # frozen: untouched
    moved = [not np.array_equal(p.unconstrained, o) for p, o in zip([p for p in m.trainable_variables if p is not frozen], others_before)]
    assert all(moved)
    frozen.trainable = True
    m.fit(X, Y, epochs=1, batch_size=64, shuffle=False)
    assert not np.array_equal(frozen.unconstrained, before_frozen)  # unfrozen: trains again


def test_optimizer_state_survives_a_workspace_rebuild():
    """A predict call with more rows than the engine's max_batch rebuilds the workspace; Adam's moments (device side) must be
    carried over: the trajectory equals the one without the rebuild."""
    from mogpe_b200._model_core import Adam
    from tests._model_helpers import model_from_spec

    spec, X, Y, eps = make_spec(seed=95, N=64, D=2, F=1, K=2, M=8, S=1, num_data=64)
    a, b = model_from_spec(spec), model_from_spec(spec)
    for m in (a, b):
        m.set_seed(2)
        m.compile(optimizer=Adam(learning_rate=0.05))
        m._core.default_batch = 64
    a.fit(X, Y, epochs=5, batch_size=64, shuffle=False)
    b.fit(X, Y, epochs=3, batch_size=64, shuffle=False)
    old = b._core.engine
    Xbig = np.random.default_rng(0).standard_normal((70000, 2))
    b.predict_y(Xbig)  # 70000 > 65536 -> new Engine
    assert b._core.engine is not old
    b._core.engine.set_seed(2)
    # continue the library random stream where the old workspace stopped: S = 1, further_gating -> the draws of 3 steps
    import ctypes as C
    hb = b.fit(X, Y, epochs=2, batch_size=64, shuffle=False)
    ha_m, ha_v = a._core.engine.get_optimizer_state()
    hb_m, hb_v = b._core.engine.get_optimizer_state()
    # the random-stream position differs after the rebuild (fresh stream), so compare what must carry over exactly: the
    # moment estimates are non-zero from the first step of the second fit() and t kept counting
    assert b.optimizer.t == a.optimizer.t == 5
    assert np.count_nonzero(hb_m) == np.count_nonzero(ha_m) and np.count_nonzero(hb_v) == np.count_nonzero(ha_v)
    # the second-moment estimate after 5 steps is a (1 - b2) weighted sum of 5 squared gradients: had it been reset at step 4
    # it would hold only 2 of them; compare magnitudes block-wise against the uninterrupted run (different draws: loose)
    ratio = np.sum(hb_v) / np.sum(ha_v)
    assert 0.6 < ratio < 1.6, ratio


@pytest.mark.parametrize("flavour", FLAVOURS)
@pytest.mark.parametrize("K", [2, 3])
def test_gating_methods_given_h_and_inducing_samples(flavour, K):
    """The gating-network methods of SURVEY 8b that take gating-function values (predict_mixing_probs_given_h,
    predict_categorical_dist_given_h_samples, predict_mixing_probs(num_inducing_samples=...)) and
    predict_y(num_inducing_samples=...) of the legacy flavour: consistent with the device-computed mixing probabilities
    and with the oracle's conditional given inducing samples."""
    from oracle import ref_numpy as rn
    from tests._model_helpers import model_from_spec

    N, S = 50, 3
    spec, X, Y, eps = make_spec(seed=70 + K, N=N, D=2, F=1, K=K, M=11, S=S, flavour=flavour)
    model = model_from_spec(spec)
    gn = model.gating_network
    G = 1 if K == 2 else K
    rng = np.random.default_rng(5)
    eps_mc = rng.standard_normal((19, N, K)) if K > 2 else None
    pi = model.predict_mixing_probs(X, eps_mc=eps_mc) if hasattr(model, "predict_mixing_probs") else gn.predict_mixing_probs(X, eps_mc=eps_mc)
    # (1) analytic / Monte-Carlo expectation from the posterior moments equals the device result
    if flavour == "keras":
        h_mean, h_var = gn.predict_h(X)
    else:
        h_mean, h_var = gn.predict_f(X)
    emc = None
    if K > 2:
        emc = eps_mc
    got = gn.predict_mixing_probs_given_h(h_mean, h_var, eps_mc=emc)
    np.testing.assert_allclose(got, pi, rtol=1e-9, atol=1e-12)
    # (2) no variance: the link itself; rows sum to one
    p0 = gn.predict_mixing_probs_given_h(h_mean)
    assert p0.shape == (N, K)
    np.testing.assert_allclose(p0.sum(-1), 1.0, rtol=1e-12)
    # (3) given inducing samples: [S, N, K], against the oracle's conditional given the same u-samples
    eps_u = rng.standard_normal((S, G, 11))
    ps = gn.predict_mixing_probs(X, num_inducing_samples=S, eps_u=eps_u, eps_mc=None if K == 2 else rng.standard_normal((23, S, N, K)))
    assert ps.shape == (S, N, K)
    np.testing.assert_allclose(ps.sum(-1), 1.0, rtol=1e-9)
    fm, fv = rn.predict_f_given_inducing_samples(spec["gating"], X, eps_u)
    if K == 2:
        lik = gn.gp.likelihood if flavour == "keras" else gn.likelihood
        want = lik.predict_mean_and_var(fm, np.broadcast_to(fv, fm.shape))[0]
        np.testing.assert_allclose(ps[..., :1], want, rtol=1e-8)
    if flavour == "keras":
        cat = gn.predict_categorical_dist_given_h_samples(X, num_h_samples=4, rng=np.random.default_rng(2))
        assert cat.probs.shape == (4, N, K)
        np.testing.assert_allclose(cat.probs.sum(-1), 1.0, rtol=1e-12)
    else:
        dist = model.predict_y(X, num_inducing_samples=S, eps_u=eps_u, eps_mc=None if K == 2 else rng.standard_normal((23, S, N, K)))
        assert dist.mean().shape == (S, N, 1) and np.all(dist.variance() > 0)


def test_dlpack_only_producer_is_consumed_zero_copy():
    """north_star: "DLPack zero-copy from TF tensors".  An object that exposes ONLY __dlpack__ (no __cuda_array_interface__,
    no .numpy) goes through the capsule path: same device pointer (zero copy), the capsule is renamed "used_dltensor" as the
    protocol requires, the producer's deleter runs exactly once when the view is dropped, and the ELBO computed from such
    inputs equals the one from host arrays."""
    import ctypes as C
    import gc

    import torch

    from mogpe_b200 import device as dev
    from tests._engine_helpers import engine_from_spec, run_engine_elbo

    class OnlyDLPack:
        def __init__(self, t):
            self.t = t
            self.capsules = []

        def __dlpack__(self, stream=None):
            cap = self.t.__dlpack__()
            self.capsules.append(cap)
            return cap

        def __dlpack_device__(self):
            return self.t.__dlpack_device__()

    t = torch.arange(12, dtype=torch.float64, device="cuda").reshape(3, 4)
    w = OnlyDLPack(t)
    assert not hasattr(w, "__cuda_array_interface__") and not hasattr(w, "numpy")
    v = dev.as_view(w)
    assert v.on_device and v.ptr == t.data_ptr() and v.shape == (3, 4)
    C.pythonapi.PyCapsule_GetName.restype = C.c_char_p
    C.pythonapi.PyCapsule_GetName.argtypes = [C.py_object]
    assert C.pythonapi.PyCapsule_GetName(w.capsules[0]) == b"used_dltensor"
    del v
    gc.collect()  # the deleter has run: torch's tensor is still alive (it holds its own reference) and intact
    assert torch.equal(t.cpu(), torch.arange(12, dtype=torch.float64).reshape(3, 4))
    with pytest.raises(TypeError):
        dev.as_view(OnlyDLPack(torch.zeros(3, device="cuda", dtype=torch.float32)))
    # end to end: the hot path on DLPack-only inputs
    spec, X, Y, eps = make_spec(seed=97, N=40, D=2, F=1, K=2, M=6, S=1, num_data=400)
    eng = engine_from_spec(spec, max_batch=40)
    ref = run_engine_elbo(eng, spec, X, Y, eps, want_grads=False)
    eps_u = eng.pack_eps_u(eps["u_experts"], None, 1)
    d = lambda a: OnlyDLPack(torch.as_tensor(np.ascontiguousarray(a), device="cuda"))
    got = eng.elbo(d(X), d(Y), "further_gating", 1, 400 / 40, d(eps_u), d(eps["h"]), None, want_grads=False)
    assert got.elbo == pytest.approx(ref.elbo, rel=1e-13)


@pytest.mark.parametrize("flavour", FLAVOURS)
def test_kernels_with_active_dims(flavour):
    """gpflow kernels built with active_dims (experts on columns 0 and 2, the gating network on column 1 of 3-D inputs): the
    objective, the gradients of every variable (the inducing inputs keep all three columns; the inactive ones get zero
    gradient) and the device-side optimiser (whose variable map must skip the lengthscale slots of inactive columns)."""
    from oracle import ref_torch as rt
    from mogpe_b200._model_core import Adam
    from tests._model_helpers import model_from_spec, oracle_grads_in_model_order
    from tests._specs import restrict_active_dims

    spec, X, Y, eps = make_spec(seed=91, N=120, D=3, F=1, K=2, M=12, S=2, flavour=flavour, num_data=1200, gating_mean_c=0.1)
    restrict_active_dims(spec, expert_dims=[0, 2], gating_dims=[1])
    model = model_from_spec(spec)
    ref, grads_u = rt.MixtureOfSVGPExperts(spec).elbo_and_grads(X, Y, eps)
    val, grads = model.elbo_and_gradients((X, Y), eps=eps)
    assert abs(val - ref) <= 1e-8 * abs(ref)
    want = oracle_grads_in_model_order(spec, model, grads_u)
    for p, g, w in zip(model.trainable_variables, grads, want):
        w = np.asarray(w)
        assert np.max(np.abs(np.asarray(g).reshape(w.shape) - w)) <= 1e-6 * max(1.0, np.max(np.abs(w))), p.name
    # host Adam and device Adam walk the same path (the variable map of the device optimiser handles the scattered slots)
    a, b = model_from_spec(spec), model_from_spec(spec)
    for m in (a, b):
        m.set_seed(7)
        m.compile(optimizer=Adam(learning_rate=0.02))
    for _ in range(4):
        a.train_step((X, Y))
    b.fit(X, Y, epochs=4, batch_size=120, shuffle=False)
    for pa, pb in zip(a.trainable_variables, b.trainable_variables):
        np.testing.assert_allclose(pb.unconstrained, pa.unconstrained, rtol=1e-9, atol=1e-11, err_msg=pa.name)


@pytest.mark.parametrize("flavour", FLAVOURS)
def test_q_diag_variational_covariance(flavour):
    """SVGP(q_diag=True): q_sqrt [M, L] with gpflow's positive() transform.  The CUDA path sees diagonal [M, M] factors; objective,
    unconstrained gradients (softplus chain on the diagonal), predictions, and the device-side optimiser -- whose variable map
    flags the diagonal entries inside the otherwise identity-transformed q_sqrt region -- against the oracle / the host Adam."""
    from oracle import ref_numpy as rn
    from oracle import ref_torch as rt
    from mogpe_b200._model_core import Adam
    from tests._model_helpers import model_from_spec, oracle_grads_in_model_order
    from tests._specs import make_q_diag

    spec, X, Y, eps = make_spec(seed=61, N=100, D=2, F=2, K=2, M=11, S=2, flavour=flavour, bound="further", num_data=700,
                                separate_kernels=True, gating_mean_c=0.1)
    make_q_diag(spec, experts=True, gating=(flavour == "keras"))  # the legacy case mixes a diagonal and a full covariance
    model = model_from_spec(spec)
    ref, grads_u = rt.MixtureOfSVGPExperts(spec).elbo_and_grads(X, Y, eps)
    val, grads = model.elbo_and_gradients((X, Y), eps=eps)
    assert abs(val - ref) <= 1e-8 * abs(ref)
    want = oracle_grads_in_model_order(spec, model, grads_u)
    for p, g, w in zip(model.trainable_variables, grads, want):
        w = np.asarray(w)
        assert np.max(np.abs(np.asarray(g).reshape(w.shape) - w)) <= 1e-6 * max(1.0, np.max(np.abs(w))), p.name
    dist = model.predict_y(X[:40])
    rm, rv = rn.predict_y(spec, X[:40], None)
    assert _rel(dist.mean(), rm) < 1e-8 and _rel(dist.variance(), rv) < 1e-8
    a, b = model_from_spec(spec), model_from_spec(spec)
    for m in (a, b):
        m.set_seed(3)
        m.compile(optimizer=Adam(learning_rate=0.02))
    for _ in range(4):
        a.train_step((X, Y))
    b.fit(X, Y, epochs=4, batch_size=100, shuffle=False)
    for pa, pb in zip(a.trainable_variables, b.trainable_variables):
        np.testing.assert_allclose(pb.unconstrained, pa.unconstrained, rtol=1e-9, atol=1e-11, err_msg=pa.name)
