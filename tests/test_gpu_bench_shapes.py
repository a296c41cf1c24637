"""GPU parity at the EXACT shapes bench.py times (BASELINE configs 2-4), against the torch-autograd oracle.

Covers the code paths the small cases never reach: kuf_kernel<4> / kuf_grad_kernel<4,1> with 16 kernel groups, K = 8 /
G = 8 Softmax gating, the split-K [M x M] reductions (C2, C3), n_rounds accumulation of Abar, the INT8-sliced products of
the backward and (when enabled) of the whole step at M = 512.

The model / minibatch generators are bench.py's own (synth_model, synth_batch), so what is checked is what is timed.
Tolerances (BASELINE.json north_star): ELBO 1e-8 relative; gradients 1e-6 relative to the largest entry of their block
AND 1e-6 in the relative L2 norm of the block (the max-norm alone would leave the small entries of a block unconstrained).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL_VALUE = 1e-8
RTOL_GRAD = 1e-6


def _workload(name):
    import bench

    wl = dict(bench.WORKLOADS[name])
    spec = bench.synth_model(wl)
    X, Y = bench.synth_batch(wl, wl["batch"], seed=100)
    K, F, M = wl["K"], wl["F"], wl["M"]
    G = 1 if K == 2 else K
    rng = np.random.default_rng(11)
    eps = {"u_experts": [rng.standard_normal((bench.S, F, M)) for _ in range(K)],
           "h": rng.standard_normal((bench.S, wl["batch"], G))}
    return wl, spec, X, Y, eps


def _compare(spec, res, ref_val, grads_u, tag):
    from tests._engine_helpers import engine_grads_named, oracle_grads_constrained

    assert np.isfinite(ref_val)
    assert abs(res.elbo - ref_val) <= RTOL_VALUE * abs(ref_val), (tag, res.elbo, ref_val)
    want = oracle_grads_constrained(spec, grads_u)
    got = engine_grads_named(res, spec)
    bad = {}
    for name, g in want.items():
        a, g = np.asarray(got[name], dtype=float), np.asarray(g, dtype=float)
        gmax = float(np.max(np.abs(g)))
        emax = float(np.max(np.abs(a - g)))
        el2 = float(np.linalg.norm((a - g).ravel()) / max(1e-300, np.linalg.norm(g.ravel())))
        if emax > RTOL_GRAD * gmax or el2 > RTOL_GRAD:
            bad[name] = (emax / max(gmax, 1e-300), el2)
    assert not bad, (tag, bad)


@pytest.fixture(scope="module")
def oracle_cache():
    return {}


def _oracle(name, cache):
    if name not in cache:
        from oracle import ref_torch as rt

        wl, spec, X, Y, eps = _workload(name)
        val, grads_u = rt.MixtureOfSVGPExperts(spec).elbo_and_grads(X, Y, eps)
        cache[name] = (wl, spec, X, Y, eps, float(val), grads_u)
    return cache[name]


@pytest.mark.parametrize("name", ["C2", "C3"])
def test_mid_size_bench_shapes_match_oracle(name, oracle_cache):
    """C2: K = 3 Softmax (3 gating kernels, shared Z), M = 128, 8192 rows.  C3: F = 2 SeparateIndependent, M = 256, 16384
    rows.  Both take the split-K path of the [M x M] reductions and the column-sliced kuf_grad."""
    from tests._engine_helpers import engine_from_spec, run_engine_elbo

    wl, spec, X, Y, eps, val, grads_u = _oracle(name, oracle_cache)
    eng = engine_from_spec(spec, max_batch=wl["batch"])
    res = run_engine_elbo(eng, spec, X, Y, eps, device_inputs=True)
    _compare(spec, res, val, grads_u, name)
    eng.close()


@pytest.mark.parametrize("mode", ["int8_step", "int8_step_8", "int8_step_7", "int8_step_6", "int8_backward", "fp64"])
def test_c4_bench_shape_matches_oracle(mode, oracle_cache):
    """C4 (the headline workload): D = 4, K = 8, G = 8 Softmax, 16 kernel groups, M = 512, 8192 rows, in the three arithmetic
    modes of a gradient step: every dense product on the INT8 tensor cores (what bench.py times: eight byte slices per operand in
    the two forward products, seven in the four gradient-only ones; eight everywhere; seven everywhere; six in the gradient-only
    products), only the two minibatch-reduced products of the backward
    there, and all-FP64 DMMA."""
    from tests._engine_helpers import engine_from_spec, run_engine_elbo

    wl, spec, X, Y, eps, val, grads_u = _oracle("C4", oracle_cache)
    eng = engine_from_spec(spec, max_batch=wl["batch"])
    eng.set_int8_step(mode.startswith("int8_step"))
    eng.set_int8_backward(mode != "fp64")
    if mode == "int8_step_8":  # byte slices loaded by the products (default: eight forward, seven gradient-only)
        eng.set_int8_grad_slices(8)
    elif mode == "int8_step_7":
        eng.set_int8_forward_slices(7)
    elif mode == "int8_step_6":  # six in the four gradient-only products
        eng.set_int8_grad_slices(6)
    eng.profile(True)
    res = run_engine_elbo(eng, spec, X, Y, eps, device_inputs=True)
    i8 = eng.profile_read_int8()["launches"]
    eng.profile(False)
    assert i8 == {"int8_step": 6, "int8_step_8": 6, "int8_step_7": 6, "int8_step_6": 6, "int8_backward": 2, "fp64": 0}[mode]  # the mode under test really ran
    _compare(spec, res, val, grads_u, f"C4 {mode}")
    # a second call on the same handle (scratch reuse, second random-stream position) reproduces the result to rounding
    res2 = run_engine_elbo(eng, spec, X, Y, eps, device_inputs=True)
    assert abs(res2.elbo - res.elbo) <= 1e-12 * abs(res.elbo)
    eng.close()


def test_c4_predict_matches_oracle(oracle_cache):
    """predict_y / mixing probabilities at the C4 model (K = 8 Softmax with explicit Monte-Carlo draws), 2048 points."""
    from oracle import ref_numpy as rn
    from tests._engine_helpers import engine_from_spec

    wl, spec, X, Y, eps, val, grads_u = _oracle("C4", oracle_cache)
    Xs = X[:2048]
    eps_mc = np.random.default_rng(5).standard_normal((13, Xs.shape[0], wl["K"]))
    eng = engine_from_spec(spec, max_batch=1024)
    out = eng.predict(Xs, eps_mc=eps_mc)
    ym, yv = rn.predict_y(spec, Xs, eps_mc)
    pi = rn.predict_mixing_probs(spec, Xs, eps_mc)
    rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    assert rel(out["y_mean"], ym) < RTOL_VALUE and rel(out["y_var"], yv) < RTOL_VALUE and rel(out["mix_probs"], pi) < RTOL_VALUE
    eng.close()


def test_c4_fast_training_mode_is_in_the_1e4_class(oracle_cache):
    """BASELINE.json's north_star allows a "1e-4" arithmetic class beside float64: with six byte slices in every product of the
    INT8 step (Engine.set_int8_forward_slices(6)) the ELBO stays within 1e-5 and every gradient block within 1e-4 of the
    oracle (measured: ~1e-7 / ~1e-6), and switching back to eight slices restores the float64-accurate step on the same handle."""
    from tests._engine_helpers import engine_from_spec, engine_grads_named, oracle_grads_constrained, run_engine_elbo

    wl, spec, X, Y, eps, val, grads_u = _oracle("C4", oracle_cache)
    eng = engine_from_spec(spec, max_batch=wl["batch"])
    eng.set_int8_forward_slices(6)
    eng.profile(True)
    res = run_engine_elbo(eng, spec, X, Y, eps, device_inputs=True)
    info = eng.profile_read_int8()
    eng.profile(False)
    assert info["launches"] == 6 and abs(info["tensor_ops"] / info["flops"] - 21.0) < 1e-9  # 21 slice pairs everywhere
    assert abs(res.elbo - val) <= 1e-5 * abs(val), (res.elbo, val)
    want = oracle_grads_constrained(spec, grads_u)
    got = engine_grads_named(res, spec)
    for name, g in want.items():
        a, g = np.asarray(got[name], dtype=float), np.asarray(g, dtype=float)
        assert np.max(np.abs(a - g)) <= 1e-4 * np.max(np.abs(g)), name
        assert np.linalg.norm((a - g).ravel()) <= 1e-4 * np.linalg.norm(g.ravel()), name
    eng.set_int8_forward_slices(8)
    res8 = run_engine_elbo(eng, spec, X, Y, eps, device_inputs=True)
    _compare(spec, res8, val, grads_u, "C4 back to eight slices")
    eng.close()
