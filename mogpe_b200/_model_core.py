"""Shared machinery of the two API flavours: SVGP parameter containers -> Engine, gradient chaining to the
unconstrained variables, light-weight distribution objects, Adam.

The reference evaluates everything through GPflow/TFP on the CPU; here every conditional, bound and
prediction goes through mogpe_b200.engine.Engine (CUDA).  What stays on the host is parameter bookkeeping
and O(N*K) post-processing of returned predictions (distribution objects for metrics / plotting callers).
"""
import math
from typing import Dict, List, Optional

import numpy as np

from .device import DeviceBuffer, PinnedBuffer, TensorView, as_view
from .engine import Engine, GPBlock, GPGrads
from .gpflow_lite import SVGP, Constant, Gaussian, Parameter

LOG_2PI = math.log(2.0 * math.pi)


def _lib_check(status):
    from . import _lib

    _lib.check(status)


def _active_dims(gp: SVGP):
    """Per latent kernel: the input columns it reads (gpflow active_dims), or None when every kernel reads all of them."""
    D = int(np.asarray(gp.Z.numpy()).shape[1])
    cols = [k.active_columns(D) if hasattr(k, "active_columns") else None for k in gp.latent_kernel_list]
    return None if all(c is None for c in cols) else cols


def svgp_to_block(gp: SVGP, is_expert: bool) -> GPBlock:
    L = gp.num_latent_gps
    mean_c = None
    if isinstance(gp.mean_function, Constant):
        mean_c = np.broadcast_to(np.asarray(gp.mean_function.c.numpy(), dtype=np.float64).reshape(-1), (L,)).copy()
    noise = None
    if is_expert:
        if not isinstance(gp.likelihood, Gaussian):
            raise NotImplementedError("experts need a Gaussian likelihood (the reference supports no other)")
        noise = np.broadcast_to(np.asarray(gp.likelihood.variance.numpy(), dtype=np.float64).reshape(-1), (L,)).copy()
    return GPBlock(
        Z=gp.Z.numpy(),
        lengthscales=[k.lengthscales.numpy() for k in gp.latent_kernel_list],
        variances=[float(k.variance.numpy()) for k in gp.latent_kernel_list],
        q_mu=gp.q_mu.numpy(),
        q_sqrt=gp.q_sqrt_full(),
        mean_c=mean_c,
        noise=noise,
        active_dims=_active_dims(gp),
    )


def accumulate_param_grads(acc: Dict[int, np.ndarray], gp: SVGP, g: GPGrads, is_expert: bool):
    """Chain constrained-space gradients to the unconstrained variables and ACCUMULATE per
    Parameter object: the reference fixture puts the same expert in the list twice (tests/keras/test_keras.py:84)."""

    def add(p: Parameter, grad_constrained):
        gu = p.chain_grad(grad_constrained)
        acc[id(p)] = acc.get(id(p), 0.0) + gu

    add(gp.Z, g.Z)
    for k, gl, gv in zip(gp.latent_kernel_list, g.lengthscales, g.variances):
        add(k.lengthscales, np.asarray(gl))
        add(k.variance, np.asarray(gv))
    add(gp.q_mu, g.q_mu)
    # (q_diag=True: only the diagonals of the [L, M, M] factors are variables -> [M, L])
    add(gp.q_sqrt, np.diagonal(g.q_sqrt, axis1=-2, axis2=-1).T if gp.q_diag else g.q_sqrt)
    if isinstance(gp.mean_function, Constant):
        c = gp.mean_function.c
        gc = g.mean_c if c.unconstrained.size == g.mean_c.size else np.sum(g.mean_c)
        add(c, np.asarray(gc))
    if is_expert:
        v = gp.likelihood.variance
        gn = g.noise if v.unconstrained.size == g.noise.size else np.sum(g.noise)
        add(v, np.asarray(gn))


class MixtureCore:
    """Owns the Engine for a list of expert SVGPs + one gating SVGP."""

    def __init__(self, expert_gps: List[SVGP], gating_gp: SVGP, flavour: str, device: int = 0):
        self.expert_gps, self.gating_gp, self.flavour, self.device = list(expert_gps), gating_gp, flavour, device
        self.engine: Optional[Engine] = None
        self.seed = 0
        self.default_batch = 1024
        self.dp = None  # (rank, world_size, nccl unique id) for data-parallel fit()
        self.precision = "float64"
        self.use_graphs = True  # fit(): replay each training step as one CUDA graph (single-GPU)

    # -- engine lifecycle ---------------------------------------------------------------------------------
    def _blocks(self):
        return [svgp_to_block(g, True) for g in self.expert_gps], svgp_to_block(self.gating_gp, False)

    def _stamp(self):
        """Identity + modification counters of every Parameter the engine's packed buffer is built from."""
        return tuple((id(p), p.version, p.trainable) for gp in self._gps() for _, p in gp.parameters())

    def ensure_engine(self, batch: int) -> Engine:
        stamp = self._stamp()
        if self.engine is not None and batch <= self.engine.max_batch and stamp == getattr(self, "_engine_stamp", None):
            return self.engine  # device parameters are already in sync with the Parameter objects
        experts, gating = self._blocks()  # (reading the Parameters first completes a deferred pull_variables)
        stamp = self._stamp()
        self._engine_stamp = stamp
        if self.engine is None or batch > self.engine.max_batch:
            if self.engine is not None:
                if getattr(self, "_opt_engine", None) is self.engine:
                    # Adam's moment estimates live on the device: carry them into the new workspace (same packed layout),
                    # otherwise the next fit() would restart them at zero while optimizer.t keeps counting
                    self._saved_opt_state = self.engine.get_optimizer_state()
                self.engine.close()
            self.engine = Engine(experts, gating, flavour=self.flavour, max_batch=max(batch, self.default_batch),
                                 device=self.device, precision=self.precision)
            self.engine.set_seed(self.seed)
            if getattr(self, "fast_training", False):
                self.engine.set_int8_forward_slices(6)
            if self.dp is not None and self.dp[1] > 1:
                self.engine.init_data_parallel(*self.dp)
        else:
            self.engine.set_params(experts, gating)
        return self.engine

    def enable_data_parallel(self, rank: int, world_size: int, unique_id: bytes):
        """One process per GPU: every rank calls fit() on ITS rows of each global minibatch; gradients are summed with
        one NCCL all-reduce per step inside the library.  unique_id: Engine.nccl_unique_id() of rank 0."""
        if self.engine is not None:
            raise RuntimeError("enable_data_parallel must be called before the first objective / fit call")
        self.dp = (int(rank), int(world_size), unique_id)

    def set_seed(self, seed: int):
        self.seed = int(seed)
        if self.engine is not None:
            self.engine.set_seed(self.seed)

    def set_precision(self, precision: str):
        """"float64" (default, gpflow's default_float); "int8" / "int8x6": predict_* run their M x M x N products on the INT8
        tensor cores with integer-sliced operands and exact accumulation (8 slices: float64-accurate, ~1.7x; 6 slices: ~1e-7,
        ~2.4x the FP64 throughput at M = 1024); "tf32": 3xTF32 on the tcgen05 tensor cores (FP32-class accuracy, ~4x).
        Gradient-free ELBO evaluations (test_step, elbo()) use the mode too; anything with gradients (fit, train_step)
        stays float64-accurate (FP64 DMMA, plus exact eight-slice INT8 products for M >= 512 and >= 2048 rows)."""
        from . import _lib

        if precision not in _lib.PRECISION_IDS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISION_IDS)}")
        if precision != self.precision:
            self.precision = precision
            if self.engine is not None:  # the workspace layout depends on the mode: rebuild on next use
                Parameter.flush_pending()
                if getattr(self, "_opt_engine", None) is self.engine:
                    self._saved_opt_state = self.engine.get_optimizer_state()
                self.engine.close()
                self.engine = None
                self._engine_stamp = self._opt_stamp = None

    def set_fast_training(self, on: bool = True):
        """Gradient steps (fit, train_step, training_loss with gradients) in the "1e-4 class" BASELINE.json allows beside
        float64: every dense product of the INT8 step loads six byte slices (42 bits) instead of eight / seven.  ELBO within
        ~1e-7 and gradients within ~1e-6 of the float64 step at the bench shape; applies when the INT8 step does (M a multiple
        of 128, M >= 512, >= 2048 rows per minibatch), smaller shapes stay on FP64 DMMA."""
        self.fast_training = bool(on)
        if self.engine is not None:
            self.engine.set_int8_forward_slices(6 if on else 8)

    @property
    def trainable_parameters(self) -> List[Parameter]:
        seen, out = set(), []
        for gp in self.expert_gps + [self.gating_gp]:
            for p in gp.trainable_parameters:
                if id(p) not in seen:
                    seen.add(id(p))
                    out.append(p)
        return out

    # -- hot path ---------------------------------------------------------------------------------------------
    def elbo(self, X, Y, bound, num_samples, num_data, eps=None, want_grads=False):
        """-> (ElboResult, {id(param): d ELBO / d unconstrained}).  eps: optional explicit draws in the oracle's
        format {"u_experts": [K][S, F, M], "u_gating": [S, G, M], "h": [S, N, G], "f": [S, N, F, K]}."""
        xv = as_view(X)
        N = int(xv.shape[0])
        eng = self.ensure_engine(N)
        scale = 1.0 if num_data is None else float(num_data) / N
        eps_u = eps_h = eps_f = None
        if eps is not None:
            if eps.get("u_experts") is not None or eps.get("u_gating") is not None:
                eps_u = eng.pack_eps_u(eps.get("u_experts"), eps.get("u_gating") if bound == "tight" else None, num_samples)
            if eps.get("h") is not None and bound != "tight":
                eps_h = np.ascontiguousarray(eps["h"], dtype=np.float64)
            if eps.get("f") is not None and bound == "further":
                eps_f = eng.pack_eps_f(eps["f"])
        if bound == "tight" and eng.G > 1:
            eng.set_mc(None if eps is None else eps.get("mc"))
        res = eng.elbo(X, Y, bound, num_samples, scale, eps_u, eps_h, eps_f, want_grads=want_grads)
        grads = None
        if want_grads:
            grads = {}
            for gp, g in zip(self.expert_gps, res.experts):
                accumulate_param_grads(grads, gp, g, True)
            accumulate_param_grads(grads, self.gating_gp, res.gating, False)
        return res, grads

    def predict(self, Xnew, want, eps_mc=None, rng=None):
        xv = as_view(Xnew)
        N = int(xv.shape[0])
        eng = self.ensure_engine(min(N, max(self.default_batch, 65536)))
        need_pi = any(w in want for w in ("mix_probs", "y_mean", "y_var"))
        if eng.G > 1 and need_pi and eps_mc is None and rng is not None:
            # gpflow Softmax.predict_mean_and_var: 100 Monte-Carlo draws from the global RNG; an explicit numpy Generator
            # reproduces that on the host, the default (rng None) draws them on the device from the library stream
            eps_mc = rng.standard_normal((100, N, eng.K))
        return eng.predict(Xnew, want=want, eps_mc=eps_mc, num_mc=100)

    # -- device-side training -------------------------------------------------------------------------
    def _gps(self):
        return self.expert_gps + [self.gating_gp]

    def _packed_blocks(self, value_of, tri_of, diag_of=None):
        """GPBlocks whose arrays are produced by value_of(param) (q_sqrt: tri_of(param) -> [L, M, M]; the [M, L] diagonal of a
        q_diag=True GP: diag_of(param), value_of by default)."""
        diag_of = diag_of or value_of
        blocks = []
        for i, gp in enumerate(self._gps()):
            is_expert = i < len(self.expert_gps)
            L = gp.num_latent_gps
            mean_c = None
            if isinstance(gp.mean_function, Constant):
                mean_c = np.broadcast_to(np.asarray(value_of(gp.mean_function.c), dtype=np.float64).reshape(-1), (L,)).copy()
            noise = None
            if is_expert:
                noise = np.broadcast_to(np.asarray(value_of(gp.likelihood.variance), dtype=np.float64).reshape(-1), (L,)).copy()
            blocks.append(GPBlock(
                Z=np.asarray(value_of(gp.Z), dtype=np.float64),
                lengthscales=[np.asarray(value_of(k.lengthscales), dtype=np.float64) for k in gp.latent_kernel_list],
                variances=[float(np.asarray(value_of(k.variance))) for k in gp.latent_kernel_list],
                q_mu=np.asarray(value_of(gp.q_mu), dtype=np.float64),
                q_sqrt=np.asarray(gp.diag_embed(diag_of(gp.q_sqrt)) if gp.q_diag else tri_of(gp.q_sqrt), dtype=np.float64),
                mean_c=mean_c, noise=noise, active_dims=_active_dims(gp)))
        return blocks[:-1], blocks[-1]

    def setup_device_optimizer(self, batch: int):
        """Build the variable map of the packed layout (ties: broadcast scalars, objects used twice) and upload the
        unconstrained variables.  -> Engine."""
        eng = self.ensure_engine(batch)
        # the variable map depends on the model structure AND on which Parameters are trainable (gpflow.set_trainable
        # staging between fit() calls, as the reference's examples do)
        flags = tuple(p.trainable for gp in self._gps() for _, p in gp.parameters())
        if getattr(self, "_opt_engine", None) is not eng or getattr(self, "_opt_flags", None) != flags:
            carried = getattr(self, "_saved_opt_state", None)
            if carried is None and getattr(self, "_opt_engine", None) is eng:
                carried = eng.get_optimizer_state()  # same workspace, only the trainable set changed
            self._saved_opt_state = None
            self._opt_stamp = None
            lowers = {float(g.likelihood.variance.transform.lower) for g in self.expert_gps}
            if len(lowers) != 1:
                raise NotImplementedError("experts with different Gaussian variance lower bounds need the host optimiser")
            ids, next_id = {}, [1]

            def ids_of(p):
                if id(p) not in ids:
                    n = p.unconstrained.size
                    ids[id(p)] = (np.arange(next_id[0], next_id[0] + n, dtype=np.float64).reshape(p.unconstrained.shape), p)
                    next_id[0] += n
                return ids[id(p)][0] if p.trainable else np.zeros(p.unconstrained.shape)

            tri = lambda p: p.transform.forward(ids_of(p))  # FillTriangular of the id vector: ids land where the values do
            flat_ids = eng.pack(*self._packed_blocks(ids_of, tri), inactive=0.0).astype(np.int64)
            uniq, first, inv = np.unique(flat_ids, return_index=True, return_inverse=True)
            rep = first[inv].astype(np.int32)
            rep[flat_ids == 0] = -1
            if any(gp.q_diag for gp in self._gps()):
                # the diagonal of a q_diag=True q_sqrt is a positive() parameter in the middle of the identity-transformed
                # q_sqrt region of the layout: flag those entries (include/mogpe_b200.h, var_rep < -1)
                zeros = lambda p: np.zeros(p.unconstrained.shape)
                mask = eng.pack(*self._packed_blocks(zeros, lambda p: np.zeros(p.numpy().shape),
                                                     lambda p: np.ones(p.unconstrained.shape)), inactive=0.0) > 0
                sel = mask & (rep >= 0)
                rep[sel] = -2 - rep[sel]
            self._var_positions = {
                pid: (first[np.searchsorted(uniq, arr.astype(np.int64).reshape(-1))].reshape(arr.shape), p)
                for pid, (arr, p) in ids.items() if p.trainable}
            eng.setup_optimizer(rep, lowers.pop())
            if carried is not None:
                eng.set_optimizer_state(*carried)  # moments of variables that stay trainable continue where they were
            self._opt_engine = eng
            self._opt_flags = flags
        if getattr(self, "_opt_stamp", None) != self._stamp():
            u_blocks = self._packed_blocks(lambda p: p.unconstrained, lambda p: p.numpy(), lambda p: p.unconstrained)
            eng.set_unconstrained(eng.pack(*u_blocks, inactive=0.0))
            self._opt_stamp = self._stamp()
        return eng

    def pull_variables(self):
        """Copy the device-side unconstrained variables back into the Parameter objects."""
        self._pull_pending = False
        eng = getattr(self, "_opt_engine", None)
        if eng is None or eng.h is None:
            return
        u = eng.u.numpy()
        for pos, p in self._var_positions.values():
            p.assign_unconstrained(u[pos])
        # the device holds exactly these variables (and params = T(u)): nothing to upload until a Parameter changes
        self._engine_stamp = self._opt_stamp = self._stamp()

    def defer_pull_variables(self):
        """fit() leaves the trained variables on the device; they are copied into the Parameter objects when a Parameter is
        next read (gpflow_lite.Parameter.flush_pending), so back-to-back fit() calls do not pay a 34 MB D2H copy each."""
        if not getattr(self, "_pull_pending", False):
            self._pull_pending = True

            def run(core=self):  # strong reference: the workspace must outlive the copy-back
                if getattr(core, "_pull_pending", False):
                    core.pull_variables()

            Parameter._pending_pulls.append(run)

    def fit_device(self, X, Y, bound, num_samples, num_data, optimizer, epochs, batch_size, shuffle, seed, tracker,
                   verbose=0, device_data=False):
        """Training loop with the whole step on the device: ELBO fwd+bwd -> bijector chain -> Adam -> params.
        Only the minibatch (H2D) and 4 result scalars (D2H) cross the bus per step."""
        import os, time
        timing = os.environ.get("MOGPE_FIT_TIMING")
        t_start = time.perf_counter()
        X, Y = np.ascontiguousarray(X, dtype=np.float64), np.ascontiguousarray(Y, dtype=np.float64)
        n = X.shape[0]
        eng = self.setup_device_optimizer(min(batch_size, n))
        t_setup = time.perf_counter()
        rng = np.random.default_rng(seed)
        history = {"loss": []}
        if device_data:
            # the data set lives in HBM; each epoch uploads one permutation, each step gathers its rows
            # on the device (no per-step host -> device copy of data)
            Xd, Yd = DeviceBuffer.from_numpy(X, self.device), DeviceBuffer.from_numpy(Y, self.device)
            idx_dev = DeviceBuffer((n,), self.device, zero=False)  # int64 bit pattern in an 8-byte slot
            bmax = min(batch_size, n)
            Xb_dev = DeviceBuffer((bmax, X.shape[1]), self.device, zero=False)
            Yb_dev = DeviceBuffer((bmax, Y.shape[1]), self.device, zero=False)
        world = 1 if self.dp is None else self.dp[1]
        steps_per_epoch = (n + batch_size - 1) // batch_size
        results = PinnedBuffer(4 * steps_per_epoch)  # every step's {elbo, var_exp, kl_g, kl_e}, copied back asynchronously
        lib = eng.lib
        for epoch in range(epochs):
            tracker.reset_states()
            order = rng.permutation(n) if shuffle else np.arange(n)
            if device_data:
                idx_dev.copy_from(np.ascontiguousarray(order, dtype=np.int64).view(np.float64))
            for step_i, s in enumerate(range(0, n, batch_size)):
                idx = order[s:s + batch_size]
                scale = 1.0 if num_data is None else float(num_data) / (len(idx) * world)
                out_ptr = results.ptr + 32 * step_i
                if world == 1:
                    # the whole step is one library call (one CUDA-graph launch once its argument set has been seen twice)
                    optimizer.t += 1
                    if device_data:
                        eng.gather_rows(Xd, idx_dev.ptr + s * 8, len(idx), X.shape[1], Xb_dev)
                        eng.gather_rows(Yd, idx_dev.ptr + s * 8, len(idx), Y.shape[1], Yb_dev)
                        Xb = TensorView(Xb_dev.ptr, (len(idx), X.shape[1]), True, self.device, Xb_dev)
                        Yb = TensorView(Yb_dev.ptr, (len(idx), Y.shape[1]), True, self.device, Yb_dev)
                    elif shuffle:
                        # the library gathers the rows straight into its pinned staging slot (no X[idx] copy on the host)
                        eng.train_step_rows(X, Y, np.ascontiguousarray(idx, dtype=np.int64), bound, num_samples, scale,
                                            optimizer.t, optimizer.lr, optimizer.b1, optimizer.b2, optimizer.eps,
                                            out_pinned_ptr=out_ptr, use_graph=self.use_graphs)
                        continue
                    else:
                        Xb, Yb = X[s:s + batch_size], Y[s:s + batch_size]  # contiguous views: no copy
                    eng.train_step(Xb, Yb, bound, num_samples, scale, optimizer.t, optimizer.lr, optimizer.b1, optimizer.b2,
                                   optimizer.eps, out_pinned_ptr=out_ptr, use_graph=self.use_graphs)
                    continue
                if device_data:
                    eng.gather_rows(Xd, idx_dev.ptr + s * 8, len(idx), X.shape[1], Xb_dev)
                    eng.gather_rows(Yd, idx_dev.ptr + s * 8, len(idx), Y.shape[1], Yb_dev)
                    Xb = TensorView(Xb_dev.ptr, (len(idx), X.shape[1]), True, self.device, Xb_dev)
                    Yb = TensorView(Yb_dev.ptr, (len(idx), Y.shape[1]), True, self.device, Yb_dev)
                    eng.elbo(Xb, Yb, bound, num_samples, scale, want_grads=True, unpack=False, readback=False)
                else:
                    # host rows -> pinned staging ring -> device, nothing waits for the GPU
                    eng.elbo_host_async(X[idx], Y[idx], bound, num_samples, scale, out_ptr)
                if world > 1:
                    eng.allreduce(eng.gbuf)  # [grads | elbo, var_exp, kl_g, kl_e] summed over ranks
                if device_data or world > 1:
                    _lib_check(lib.mogpe_memcpy_d2h_async(eng.device, out_ptr, eng.out_ptr, 32, None))
                optimizer.t += 1
                eng.adam_step(optimizer.t, optimizer.lr, optimizer.b1, optimizer.b2, optimizer.eps)
            eng.check_numerics()  # one synchronisation per epoch; raises if a Cholesky failed
            losses = -results.array[0:4 * steps_per_epoch:4]
            for v in losses:
                tracker.update_state(v)
            history["loss"].append(tracker.result())
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs} - loss: {tracker.result():.4f}")
        t_loop = time.perf_counter()
        self.defer_pull_variables()
        if timing:
            print(f"fit timing: setup {1e3 * (t_setup - t_start):.1f} ms, loop {1e3 * (t_loop - t_setup):.1f} ms, "
                  f"pull {1e3 * (time.perf_counter() - t_loop):.1f} ms", flush=True)
        return history

    def latent_slices(self):
        """-> per GP (experts..., gating): slice into the latent axis of f_mean / f_var."""
        out, l = [], 0
        for gp in self.expert_gps + [self.gating_gp]:
            out.append(slice(l, l + gp.num_latent_gps))
            l += gp.num_latent_gps
        return out


# ---- distributions returned to callers (metrics / plotting) -------------------------------------------------------
class Normal:
    """Stand-in for tfd.Normal / tfd.MultivariateNormalDiag(loc, scale) with the methods the reference's callers
    use.  `event_ndims` = 1 reduces log_prob over the last axis (MultivariateNormalDiag)."""

    def __init__(self, loc, scale, event_ndims=0):
        self.loc, self.scale, self.event_ndims = np.asarray(loc), np.asarray(scale), event_ndims

    def mean(self):
        return self.loc

    def stddev(self):
        return np.abs(self.scale)

    def variance(self):
        return self.scale**2

    def log_prob(self, y):
        z = (np.asarray(y) - self.loc) / self.scale
        lp = -0.5 * z * z - np.log(np.abs(self.scale)) - 0.5 * LOG_2PI
        return lp.sum(-1) if self.event_ndims else lp

    def prob(self, y):
        return np.exp(self.log_prob(y))

    def sample(self, n=None, rng=None):
        rng = rng or np.random.default_rng()
        shape = self.loc.shape if n is None else (n,) + self.loc.shape
        return self.loc + self.scale * rng.standard_normal(shape)

    @property
    def batch_shape(self):
        return self.loc.shape[:-1] if self.event_ndims else self.loc.shape


class Categorical:
    def __init__(self, probs):
        self.probs = np.asarray(probs)

    def probs_parameter(self):
        return self.probs

    def sample(self, n=None, rng=None):
        rng = rng or np.random.default_rng()
        c = np.cumsum(self.probs, -1)
        u = rng.random(self.probs.shape[:-1] if n is None else (n,) + self.probs.shape[:-1])
        return (u[..., None] > c).sum(-1)


class MixtureDistribution:
    """Stand-in for the tfd.Mixture (Keras flavour) / tfd.MixtureSameFamily (legacy) returned by predict_y.
    mean() and variance() are the device-computed values (flavour-exact)."""

    def __init__(self, mixing_probs, locs, scales, mean, variance):
        self.mixing_probs = np.asarray(mixing_probs)  # [N, K]
        self.locs, self.scales = np.asarray(locs), np.asarray(scales)  # [N, F, K]
        self._mean, self._variance = np.asarray(mean), np.asarray(variance)  # [N, F]

    def mean(self):
        return self._mean

    def variance(self):
        return self._variance

    def stddev(self):
        return np.sqrt(self._variance)

    def log_prob(self, y, joint_over_outputs=True):
        """Keras flavour (joint_over_outputs): log sum_k pi_k prod_f N(y_f; mu_kf, s_kf) -> [N];
        legacy MixtureSameFamily over Normal: per output -> [N, F]."""
        y = np.asarray(y)[..., None]
        z = (y - self.locs) / self.scales
        lp = -0.5 * z * z - np.log(np.abs(self.scales)) - 0.5 * LOG_2PI  # [N, F, K]
        if joint_over_outputs:
            w = lp.sum(-2) + np.log(self.mixing_probs)
        else:
            w = lp + np.log(self.mixing_probs)[..., None, :]
        m = w.max(-1, keepdims=True)
        return (m + np.log(np.exp(w - m).sum(-1, keepdims=True)))[..., 0]

    def prob(self, y, **kw):
        return np.exp(self.log_prob(y, **kw))

    def sample(self, n=1, rng=None):
        rng = rng or np.random.default_rng()
        N, F, K = self.locs.shape
        k = Categorical(self.mixing_probs).sample(n, rng)  # [n, N]
        idx = np.broadcast_to(k[..., None, None], (n, N, F, 1))
        loc = np.take_along_axis(np.broadcast_to(self.locs, (n, N, F, K)), idx, -1)[..., 0]
        sc = np.take_along_axis(np.broadcast_to(self.scales, (n, N, F, K)), idx, -1)[..., 0]
        return loc + sc * rng.standard_normal((n, N, F))


# ---- optimiser (the reference uses tf.optimizers.Adam everywhere) ----------------------------
class Adam:
    """TF-default Adam (beta_1 = 0.9, beta_2 = 0.999, epsilon = 1e-7) over Parameter.unconstrained arrays."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.t = 0
        self.m: Dict[int, np.ndarray] = {}
        self.v: Dict[int, np.ndarray] = {}

    def apply_gradients(self, grads_and_vars):
        """grads of the LOSS w.r.t. the unconstrained variables, as tf.optimizers.Adam.apply_gradients."""
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2**self.t) / (1.0 - self.b1**self.t)
        for g, p in grads_and_vars:
            if g is None:
                continue
            g = np.asarray(g, dtype=np.float64).reshape(p.unconstrained.shape)
            m = self.m.get(id(p))
            if m is None:
                m = self.m[id(p)] = np.zeros_like(p.unconstrained)
                self.v[id(p)] = np.zeros_like(p.unconstrained)
            v = self.v[id(p)]
            m += (1.0 - self.b1) * (g - m)
            v += (1.0 - self.b2) * (g * g - v)
            p.assign_unconstrained(p.unconstrained - lr_t * m / (np.sqrt(v) + self.eps))


class MeanTracker:
    """tf.keras.metrics.Mean."""

    def __init__(self, name="loss"):
        self.name, self.total, self.count = name, 0.0, 0

    def update_state(self, v):
        self.total += float(v)
        self.count += 1

    def result(self):
        return self.total / max(1, self.count)

    def reset_states(self):
        self.total, self.count = 0.0, 0
