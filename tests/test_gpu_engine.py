"""GPU parity tests of the CUDA library (through the C-ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): ELBO / predictions 1e-8 relative, gradients 1e-6 relative (float64).
"""
import ctypes as C

import numpy as np
import pytest

from tests._specs import BOUNDS, FLAVOURS, make_spec

pytestmark = pytest.mark.gpu

RTOL_VALUE = 1e-8
RTOL_GRAD = 1e-6


def _rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


@pytest.mark.parametrize("cfg", [-1, 0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(64, 64, 64), (128, 256, 192), (256, 1024, 128), (512, 4096, 512)])
def test_dmma_gemm_matches_numpy(ta, tb, M, N, K, cfg):
    import torch
    from mogpe_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(M + N + K + ta * 2 + tb)
    batch = 3
    A = rng.standard_normal((batch, K, M) if ta else (batch, M, K))
    B = rng.standard_normal((batch, N, K) if tb else (batch, K, N))
    C0 = rng.standard_normal((batch, M, N))
    dA, dB, dC = (torch.as_tensor(x, device="cuda") for x in (A, B, C0))
    _lib.check(lib.mogpe_debug_gemm(0, dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), batch, M, N, K, ta, tb, 0, 0, 1.5, 0.5, cfg, None))
    torch.cuda.synchronize()
    opA = np.swapaxes(A, 1, 2) if ta else A
    opB = np.swapaxes(B, 1, 2) if tb else B
    ref = 1.5 * opA @ opB + 0.5 * C0
    assert _rel(dC.cpu().numpy(), ref) < 1e-13


@pytest.mark.parametrize("cfg", [-1, 0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("size", [128, 256, 512])
def test_dmma_gemm_triangular_modes(size, cfg):
    import torch
    from mogpe_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(size)
    M = size
    T = np.tril(rng.standard_normal((2, M, M)))
    R = rng.standard_normal((2, M, 1024))
    out = torch.zeros((2, M, 1024), dtype=torch.float64, device="cuda")
    dT, dR = torch.as_tensor(T, device="cuda"), torch.as_tensor(R, device="cuda")
    # left lower: T @ R
    _lib.check(lib.mogpe_debug_gemm(0, dT.data_ptr(), dR.data_ptr(), out.data_ptr(), 2, M, 1024, M, 0, 0, 1, 0, 1.0, 0.0, cfg, None))
    assert _rel(out.cpu().numpy(), T @ R) < 1e-13
    # left upper (transposed lower): T^T @ R
    _lib.check(lib.mogpe_debug_gemm(0, dT.data_ptr(), dR.data_ptr(), out.data_ptr(), 2, M, 1024, M, 1, 0, 2, 0, 1.0, 0.0, cfg, None))
    assert _rel(out.cpu().numpy(), np.swapaxes(T, 1, 2) @ R) < 1e-13
    # tril output of an NT product over a long k
    W = rng.standard_normal((2, M, 2048))
    A = rng.standard_normal((2, M, 2048))
    o2 = torch.full((2, M, M), 7.0, dtype=torch.float64, device="cuda")
    dW, dA = torch.as_tensor(W, device="cuda"), torch.as_tensor(A, device="cuda")
    _lib.check(lib.mogpe_debug_gemm(0, dW.data_ptr(), dA.data_ptr(), o2.data_ptr(), 2, M, M, 2048, 0, 1, 0, 1, -1.0, 0.0, cfg, None))
    assert _rel(o2.cpu().numpy(), -np.tril(W @ np.swapaxes(A, 1, 2))) < 1e-13
    # right lower: R2 @ T  and left-upper x right-lower: T^T @ T2
    R2 = rng.standard_normal((2, M, M))
    o3 = torch.zeros((2, M, M), dtype=torch.float64, device="cuda")
    dR2 = torch.as_tensor(R2, device="cuda")
    _lib.check(lib.mogpe_debug_gemm(0, dR2.data_ptr(), dT.data_ptr(), o3.data_ptr(), 2, M, M, M, 0, 0, 4, 0, 1.0, 0.0, cfg, None))
    assert _rel(o3.cpu().numpy(), R2 @ T) < 1e-13
    _lib.check(lib.mogpe_debug_gemm(0, dT.data_ptr(), dT.data_ptr(), o3.data_ptr(), 2, M, M, M, 1, 0, 2 | 4, 0, 1.0, 0.0, cfg, None))
    assert _rel(o3.cpu().numpy(), np.swapaxes(T, 1, 2) @ T) < 1e-13


def _check_intermediates(eng, spec, X):
    """Kuu, L, Linv, Kuf, A of every group against numpy (A via solve_triangular, as the reference does)."""
    from scipy.linalg import solve_triangular

    from oracle import ref_numpy as rn

    Mp, Bp = eng.Mp, ((eng.max_batch + 127) // 128) * 128
    Q = eng.Q
    Kuu = eng.debug("Kuu").reshape(Q, Mp, Mp)
    L = eng.debug("L").reshape(Q, Mp, Mp)
    Linv = eng.debug("Linv").reshape(Q, Mp, Mp)
    Kuf = eng.debug("Kuf").reshape(Q, Mp, Bp)
    A = eng.debug("A").reshape(Q, Mp, Bp)
    N = X.shape[0]
    g = 0
    worst = {}
    for gp in spec["experts"] + [spec["gating"]]:
        Z = np.asarray(gp["Z"])
        M = Z.shape[0]
        for kern in gp["kernels"]:
            K_ref = rn.rbf_K(kern, Z) + 1e-6 * np.eye(M)
            L_ref = np.linalg.cholesky(K_ref)
            Kuf_ref = rn.rbf_K(kern, Z, X)
            A_ref = solve_triangular(L_ref, Kuf_ref, lower=True)
            for name, got, ref in (("Kuu", Kuu[g, :M, :M], K_ref), ("L", L[g, :M, :M], L_ref),
                                   ("Linv", Linv[g, :M, :M], np.linalg.inv(L_ref)),
                                   ("Kuf", Kuf[g, :M, :N], Kuf_ref), ("A", A[g, :M, :N], A_ref)):
                worst[name] = max(worst.get(name, 0.0), _rel(got, ref))
            # padding: identity / zeros
            assert np.array_equal(L[g, M:, M:], np.eye(Mp - M))
            assert not np.any(L[g, M:, :M]) and not np.any(Kuf[g, M:, :]) and not np.any(A[g, M:, :N])
            assert not np.any(np.triu(L[g], 1)) and not np.any(np.triu(Linv[g], 1))
            g += 1
    return worst


@pytest.mark.parametrize("M,N", [(8, 24), (70, 300), (200, 1000)])
def test_factorisation_and_conditional_intermediates(M, N):
    from tests._engine_helpers import engine_from_spec, run_engine_elbo

    spec, X, Y, eps = make_spec(seed=M, N=N, D=2, F=1, K=2, M=M, separate_kernels=False)
    eng = engine_from_spec(spec, max_batch=N)
    run_engine_elbo(eng, spec, X, Y, eps, want_grads=False)
    worst = _check_intermediates(eng, spec, X)
    assert worst["Kuu"] < 1e-14 and worst["Kuf"] < 1e-13, worst
    # conditioning of Kuu (jitter 1e-6) amplifies rounding in L / L^-1 / A; 1e-9 is far inside the ELBO budget
    assert worst["L"] < 1e-9 and worst["A"] < 1e-8, worst


CASES = [
    # K, F, S, M, N, D, separate kernels
    (2, 1, 1, 8, 24, 2, False),
    (2, 2, 2, 20, 100, 3, True),
    (3, 1, 2, 33, 130, 1, False),
    (3, 2, 3, 70, 257, 2, True),
]


@pytest.mark.parametrize("flavour", FLAVOURS)
@pytest.mark.parametrize("bound", BOUNDS)
@pytest.mark.parametrize("case", CASES)
def test_elbo_and_gradients_match_oracle(flavour, bound, case):
    from oracle import ref_numpy as rn
    from oracle import ref_torch as rt
    from tests._engine_helpers import engine_from_spec, engine_grads_named, oracle_grads_constrained, run_engine_elbo

    K, F, S, M, N, D, sep = case
    spec, X, Y, eps = make_spec(seed=100 + K + F + S, N=N, D=D, F=F, K=K, M=M, S=S, flavour=flavour, bound=bound,
                                num_data=10 * N, separate_kernels=sep, gating_mean_c=0.2)
    if bound == "tight" and K > 2:  # Softmax predict_mean_and_var is Monte-Carlo: explicit draws [S, R, N, K]
        eps["mc"] = np.random.default_rng(17).standard_normal((S, 9, N, K))
    eng = engine_from_spec(spec, max_batch=N)
    res = run_engine_elbo(eng, spec, X, Y, eps)
    ref = rn.elbo(spec, X, Y, eps)
    assert np.isfinite(ref)
    assert abs(res.elbo - ref) <= RTOL_VALUE * abs(ref), (res.elbo, ref)
    assert abs(res.elbo - (res.var_exp - res.kl_gating - res.kl_experts)) <= 1e-12 * abs(ref)
    val, grads_u = rt.MixtureOfSVGPExperts(spec).elbo_and_grads(X, Y, eps)
    want = oracle_grads_constrained(spec, grads_u)
    got = engine_grads_named(res, spec)
    bad = {}
    for name, g_ref in want.items():
        r = _rel(got[name], g_ref)
        g_ref = np.asarray(g_ref, dtype=float)
        diff = np.asarray(got[name], dtype=float) - g_ref
        scale = max(1.0, float(np.max(np.abs(g_ref))))
        # max-norm against the block's largest entry AND the relative L2 norm of the whole block (the max-norm alone leaves
        # the small entries of a large block unconstrained); blocks that are numerically zero are held to the absolute bound
        l2 = float(np.linalg.norm(diff.ravel()) / max(np.linalg.norm(g_ref.ravel()), 1.0))
        if np.max(np.abs(diff)) > RTOL_GRAD * scale or l2 > RTOL_GRAD:
            bad[name] = (r, l2)
    assert not bad, bad


@pytest.mark.parametrize("device_inputs", [False, True])
def test_host_and_device_entry_points_agree(device_inputs):
    from tests._engine_helpers import engine_from_spec, run_engine_elbo

    spec, X, Y, eps = make_spec(seed=5, N=200, D=2, F=1, K=2, M=16, S=2)
    eng = engine_from_spec(spec, max_batch=256)
    a = run_engine_elbo(eng, spec, X, Y, eps, device_inputs=device_inputs)
    b = run_engine_elbo(eng, spec, X, Y, eps, device_inputs=not device_inputs)
    assert a.elbo == pytest.approx(b.elbo, rel=1e-13)
    np.testing.assert_allclose(a.gating.q_sqrt, b.gating.q_sqrt, rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("flavour", FLAVOURS)
@pytest.mark.parametrize("K,F,sep", [(2, 1, False), (2, 2, True), (3, 2, True)])
def test_predict_matches_oracle(flavour, K, F, sep):
    from oracle import ref_numpy as rn
    from tests._engine_helpers import engine_from_spec

    N = 333
    spec, X, Y, eps = make_spec(seed=9 + K, N=N, D=2, F=F, K=K, M=25, flavour=flavour, separate_kernels=sep)
    eng = engine_from_spec(spec, max_batch=128)  # forces 3 chunks
    eps_mc = np.random.default_rng(3).standard_normal((17, N, K)) if K > 2 else None
    out = eng.predict(X, eps_mc=eps_mc)
    pi = rn.predict_mixing_probs(spec, X, eps_mc)
    ym, yv = rn.predict_y(spec, X, eps_mc)
    assert _rel(out["mix_probs"], pi) < RTOL_VALUE
    assert _rel(out["y_mean"], ym) < RTOL_VALUE
    assert _rel(out["y_var"], yv) < RTOL_VALUE
    l = 0
    for gp in spec["experts"] + [spec["gating"]]:
        fm, fv = rn.predict_f(gp, X)
        L = fm.shape[1]
        assert _rel(out["f_mean"][:, l:l + L], fm) < RTOL_VALUE
        assert _rel(out["f_var"][:, l:l + L], fv) < RTOL_VALUE
        l += L


def test_different_inducing_counts_per_gp():
    from oracle import ref_numpy as rn
    from tests._engine_helpers import engine_from_spec, run_engine_elbo

    spec, X, Y, eps = make_spec(seed=21, N=150, D=1, F=1, K=2, M=12, M_experts=[6, 7], M_gating=40, S=2, num_data=1000)
    eng = engine_from_spec(spec, max_batch=150)
    res = run_engine_elbo(eng, spec, X, Y, eps)
    ref = rn.elbo(spec, X, Y, eps)
    assert abs(res.elbo - ref) <= RTOL_VALUE * abs(ref)


def test_fill_normal_is_standard_normal_and_reproducible():
    from mogpe_b200.device import DeviceBuffer
    from tests._engine_helpers import engine_from_spec

    spec, X, Y, eps = make_spec(seed=1, N=16, M=4)
    eng = engine_from_spec(spec, max_batch=16)
    buf = DeviceBuffer((200001,), zero=True)
    eng.fill_normal(buf, seed=1234, offset=0)
    a = buf.numpy()
    eng.fill_normal(buf, seed=1234, offset=0)
    assert np.array_equal(a, buf.numpy())
    assert abs(a.mean()) < 0.01 and abs(a.std() - 1.0) < 0.01
    from scipy import stats

    assert stats.kstest(a, "norm").pvalue > 1e-4
    # counter-based: a window at an even offset reproduces the same stream
    eng.fill_normal(buf, seed=1234, offset=1000, count=500)
    assert np.array_equal(buf.numpy()[:500], a[1000:1500])


def test_invalid_arguments_raise():
    from mogpe_b200._lib import MogpeError
    from tests._engine_helpers import engine_from_spec, run_engine_elbo

    spec, X, Y, eps = make_spec(seed=2, N=32, M=4, K=3)
    eng = engine_from_spec(spec, max_batch=16)
    with pytest.raises(MogpeError):
        run_engine_elbo(eng, spec, X, Y, eps)  # N > max_batch
    eng = engine_from_spec(spec, max_batch=32)
    with pytest.raises(MogpeError, match="Softmax"):
        eng.predict(X, want=("mix_probs",), num_mc=0)  # Softmax mixing probabilities need Monte-Carlo draws
    with pytest.raises(ValueError):
        eng.elbo(X[:, :1].repeat(3, 1), Y)


def test_large_inducing_count_m1024_predict_and_elbo():
    """BASELINE config C5 uses M = 1024 (Mp = 1024: 32 Cholesky block steps, 16 GEMM row blocks)."""
    from oracle import ref_numpy as rn
    from tests._engine_helpers import engine_from_spec, run_engine_elbo

    spec, X, Y, eps = make_spec(seed=31, N=1500, D=2, F=1, K=2, M=1024, S=1, num_data=15000)
    eng = engine_from_spec(spec, max_batch=512)
    out = eng.predict(X)  # 3 chunks
    ym, yv = rn.predict_y(spec, X)
    assert _rel(out["y_mean"], ym) < RTOL_VALUE and _rel(out["y_var"], yv) < RTOL_VALUE
    assert _rel(out["mix_probs"], rn.predict_mixing_probs(spec, X)) < RTOL_VALUE
    Xb, Yb = X[:512], Y[:512]
    eps_b = {"u_experts": eps["u_experts"], "h": eps["h"][:, :512]}
    res = run_engine_elbo(eng, spec, Xb, Yb, eps_b, want_grads=False)
    ref = rn.elbo(spec, Xb, Yb, eps_b)
    assert abs(res.elbo - ref) <= RTOL_VALUE * abs(ref)


def test_cholesky_failure_is_reported_like_the_reference():
    """tf.linalg.cholesky raises for a non positive-definite Kuu; we report MOGPE_ERR_NUMERIC naming the group."""
    from mogpe_b200._lib import MogpeError
    from tests._engine_helpers import engine_from_spec, run_engine_elbo

    spec, X, Y, eps = make_spec(seed=3, N=40, D=1, K=2, M=6)
    spec["gating"]["kernels"][0]["variance"] = float("nan")
    eng = engine_from_spec(spec, max_batch=40)
    with pytest.raises(MogpeError, match="Cholesky decomposition was not successful.*group 2"):
        run_engine_elbo(eng, spec, X, Y, eps)
    with pytest.raises(MogpeError, match="Cholesky"):
        run_engine_elbo(eng, spec, X, Y, eps, device_inputs=True)
    spec["gating"]["kernels"][0]["variance"] = 1.0
    eng2 = engine_from_spec(spec, max_batch=40)
    assert np.isfinite(run_engine_elbo(eng2, spec, X, Y, eps).elbo)


def test_predict_softmax_library_draws_are_chunking_independent_and_close_to_explicit_draws():
    """eps_mc=None: the Softmax Monte-Carlo draws come from the library stream, indexed by the global point index, so the
    result does not depend on max_batch; with many draws it converges to the explicit-draw estimate (both are unbiased
    estimates of the same expectation)."""
    from tests._engine_helpers import engine_from_spec

    N, K = 500, 3
    spec, X, Y, eps = make_spec(seed=31, N=N, D=2, F=1, K=K, M=20)
    a, b = engine_from_spec(spec, max_batch=128), engine_from_spec(spec, max_batch=512)
    for e in (a, b):
        e.set_seed(7)
    pa = a.predict(X, want=("mix_probs", "y_mean"), num_mc=64)
    pb = b.predict(X, want=("mix_probs", "y_mean"), num_mc=64)
    np.testing.assert_allclose(pa["mix_probs"], pb["mix_probs"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(pa["mix_probs"].sum(-1), 1.0, rtol=1e-12)
    pa2 = a.predict(X, want=("mix_probs",), num_mc=64)  # the stream advances between calls
    assert not np.array_equal(pa2["mix_probs"], pa["mix_probs"])
    big = a.predict(X, want=("mix_probs",), num_mc=4000)["mix_probs"]
    ref = a.predict(X, want=("mix_probs",), eps_mc=np.random.default_rng(0).standard_normal((4000, N, K)))["mix_probs"]
    assert np.max(np.abs(big - ref)) < 0.03  # two independent 4000-draw estimates: std <= 0.5 / sqrt(4000) each


@pytest.mark.parametrize("K,F,S,sep", [(3, 1, 2, True), (2, 2, 1, False)])
@pytest.mark.parametrize("bound", BOUNDS)
def test_int8_step_general_structures_match_oracle(bound, K, F, S, sep):
    """The INT8-sliced step on model structures the bench never sees: several samples, a shared kernel with two latents per
    group (two marginal latents accumulate into one Abar), Softmax gating with separate kernels, every bound (tight: no
    marginal latent at all; further: the experts are marginal too), inducing counts below the padded size."""
    from oracle import ref_torch as rt
    from tests._engine_helpers import engine_from_spec, engine_grads_named, oracle_grads_constrained, run_engine_elbo

    N = 2100  # not a multiple of 64: ragged last column tile
    spec, X, Y, eps = make_spec(seed=23 + K, N=N, D=3, F=F, K=K, M=400, S=S, flavour="keras", bound=bound, num_data=30000,
                                separate_kernels=sep, M_gating=512, gating_mean_c=0.1)
    if bound == "tight" and K > 2:
        eps["mc"] = np.random.default_rng(3).standard_normal((S, 7, N, K))
    eng = engine_from_spec(spec, max_batch=N)
    ref, grads_u = rt.MixtureOfSVGPExperts(spec).elbo_and_grads(X, Y, eps)
    want = oracle_grads_constrained(spec, grads_u)
    eng.profile(True)
    res = run_engine_elbo(eng, spec, X, Y, eps)
    assert eng.profile_read_int8()["launches"] >= 3
    eng.profile(False)
    assert abs(res.elbo - ref) <= RTOL_VALUE * abs(ref), (res.elbo, ref)
    got = engine_grads_named(res, spec)
    bad = {}
    for name, g in want.items():
        err = np.max(np.abs(np.asarray(got[name]) - g))
        if err > RTOL_GRAD * max(1.0, np.max(np.abs(g))):
            bad[name] = err / max(1.0, np.max(np.abs(g)))
    assert not bad, bad
    eng.close()


def test_int8_step_with_active_dims():
    """Kernels restricted to a subset of the input columns (gpflow active_dims) on the INT8 path: the inactive columns carry
    the huge lengthscale of engine.INACTIVE_LENGTHSCALE through kuf8 / xs_prep / kuf_grad_slice / kuu."""
    from oracle import ref_torch as rt
    from tests._engine_helpers import engine_from_spec, engine_grads_named, oracle_grads_constrained, run_engine_elbo
    from tests._specs import restrict_active_dims

    N = 2048
    spec, X, Y, eps = make_spec(seed=55, N=N, D=3, F=1, K=2, M=512, S=1, flavour="keras", bound="further_gating", num_data=30000)
    restrict_active_dims(spec, expert_dims=[0, 2], gating_dims=[1, 2])
    eng = engine_from_spec(spec, max_batch=N)
    ref, grads_u = rt.MixtureOfSVGPExperts(spec).elbo_and_grads(X, Y, eps)
    want = oracle_grads_constrained(spec, grads_u)
    eng.profile(True)
    res = run_engine_elbo(eng, spec, X, Y, eps)
    assert eng.profile_read_int8()["launches"] == 6
    eng.profile(False)
    assert abs(res.elbo - ref) <= RTOL_VALUE * abs(ref), (res.elbo, ref)
    got = engine_grads_named(res, spec)
    bad = {}
    for name, g in want.items():
        err = np.max(np.abs(np.asarray(got[name]) - g))
        if err > RTOL_GRAD * max(1.0, np.max(np.abs(g))):
            bad[name] = err / max(1.0, np.max(np.abs(g)))
    assert not bad, bad
    eng.close()


def test_int8_step_applies_above_512_inducing_points():
    """M = 520 is padded to 640 (a multiple of the 128-row tile), so the step runs on the INT8 tensor cores; the 120 padding
    rows / columns of every operand must stay inert (unit diagonal in Kuu, zero elsewhere)."""
    from oracle import ref_torch as rt
    from tests._engine_helpers import engine_from_spec, engine_grads_named, oracle_grads_constrained, run_engine_elbo

    N = 2048
    spec, X, Y, eps = make_spec(seed=77, N=N, D=2, F=1, K=2, M=520, S=1, flavour="keras", bound="further_gating", num_data=40000)
    eng = engine_from_spec(spec, max_batch=N)
    assert eng.Mp == 640
    ref, grads_u = rt.MixtureOfSVGPExperts(spec).elbo_and_grads(X, Y, eps)
    want = oracle_grads_constrained(spec, grads_u)
    eng.profile(True)
    res = run_engine_elbo(eng, spec, X, Y, eps)
    assert eng.profile_read_int8()["launches"] == 6
    eng.profile(False)
    assert abs(res.elbo - ref) <= RTOL_VALUE * abs(ref), (res.elbo, ref)
    got = engine_grads_named(res, spec)
    bad = {}
    for name, g in want.items():
        err = np.max(np.abs(np.asarray(got[name]) - g))
        if err > RTOL_GRAD * max(1.0, np.max(np.abs(g))):
            bad[name] = err / max(1.0, np.max(np.abs(g)))
    assert not bad, bad
    eng.close()


@pytest.mark.parametrize("D", [1, 6])
def test_int8_step_input_dimensions_match_oracle(D):
    """The INT8 step for input dimensions on either side of the bench's D = 4: D = 1 (kuf8_kernel<2>, padded 4-wide records of
    the pre-scaled inputs) and D = 6 (the 16-wide instantiations kuf8_kernel<16>, xs_prep_kernel<16>, kuf_grad_slice_kernel<16, 8>)."""
    from oracle import ref_torch as rt
    from tests._engine_helpers import engine_from_spec, engine_grads_named, oracle_grads_constrained, run_engine_elbo

    N = 2048
    spec, X, Y, eps = make_spec(seed=40 + D, N=N, D=D, F=1, K=2, M=512, S=1, flavour="keras", bound="further_gating",
                                num_data=50000)
    eng = engine_from_spec(spec, max_batch=N)
    ref, grads_u = rt.MixtureOfSVGPExperts(spec).elbo_and_grads(X, Y, eps)
    want = oracle_grads_constrained(spec, grads_u)
    eng.profile(True)
    res = run_engine_elbo(eng, spec, X, Y, eps)
    assert eng.profile_read_int8()["launches"] == 6
    eng.profile(False)
    assert abs(res.elbo - ref) <= RTOL_VALUE * abs(ref), (res.elbo, ref)
    got = engine_grads_named(res, spec)
    bad = {}
    for name, g in want.items():
        err = np.max(np.abs(np.asarray(got[name]) - g))
        if err > RTOL_GRAD * max(1.0, np.max(np.abs(g))):
            bad[name] = err / max(1.0, np.max(np.abs(g)))
    assert not bad, bad
    eng.close()


@pytest.mark.parametrize("flavour", FLAVOURS)
@pytest.mark.parametrize("bound", ["further_gating", "further"])
def test_int8_backward_products_match_oracle_gradients(flavour, bound):
    """At M >= 512 (multiple of 128) and >= 2048 rows the two minibatch-reduced [M x M] results of the backward (Lbar, Sbar) run
    on the INT8 tensor cores (integer-sliced operands, exact accumulation).  Gradients must still meet the float64 tolerance
    against the oracle, agree with the FP64 DMMA path, and really come from a different code path."""
    from oracle import ref_torch as rt
    from tests._engine_helpers import engine_from_spec, engine_grads_named, oracle_grads_constrained, run_engine_elbo

    spec, X, Y, eps = make_spec(seed=17, N=2048, D=2, F=1, K=2, M=512, S=1, flavour=flavour, bound=bound, num_data=20000)
    eng = engine_from_spec(spec, max_batch=2048)
    ref, grads_u = rt.MixtureOfSVGPExperts(spec).elbo_and_grads(X, Y, eps)
    want = oracle_grads_constrained(spec, grads_u)
    res = run_engine_elbo(eng, spec, X, Y, eps)
    assert abs(res.elbo - ref) <= RTOL_VALUE * abs(ref)
    got = engine_grads_named(res, spec)
    for name, g in want.items():
        assert np.max(np.abs(np.asarray(got[name]) - g)) <= RTOL_GRAD * max(1.0, np.max(np.abs(g))), name
    eng.set_int8_step(False)  # hybrid: only Lbar / Sbar on the INT8 tensor cores
    res_h = run_engine_elbo(eng, spec, X, Y, eps)
    assert abs(res_h.elbo - ref) <= RTOL_VALUE * abs(ref)
    got_h = engine_grads_named(res_h, spec)
    for name, g in want.items():
        assert np.max(np.abs(np.asarray(got_h[name]) - g)) <= RTOL_GRAD * max(1.0, np.max(np.abs(g))), name
    eng.set_int8_backward(False)
    got64 = engine_grads_named(run_engine_elbo(eng, spec, X, Y, eps), spec)
    differs = False
    for name, g in got64.items():
        a, b = np.asarray(got[name]), np.asarray(g)
        # Lbar carries ~1e-12 (8 byte slices, K = 2048) and the Cholesky backward of this ill-conditioned Kuu (512 inducing
        # points in 2-D) amplifies it to ~1e-9 of the largest gradient entry: still 3 orders inside the 1e-6 tolerance
        assert np.max(np.abs(a - b)) <= 1e-7 * max(1.0, np.max(np.abs(b))), name
        differs = differs or not np.array_equal(a, b)
    assert differs
    eng.close()
