"""Host-side engine: packs the parameters of K SVGP experts + one SVGP gating network into the layout of
include/mogpe_b200.h and drives the CUDA library.  Everything numerical happens on the device; this
module only moves parameters / gradients between the model objects and the packed buffers.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _lib
from .device import DeviceBuffer, as_view

DEFAULT_JITTER = 1e-6  # gpflow.config.default_jitter()
# lengthscale of the input columns outside a kernel's active_dims: x / l and z / l vanish, so does every gradient through them
INACTIVE_LENGTHSCALE = 1e150


@dataclass
class GPBlock:
    """Constrained-space parameters of one SVGP: shared Z, one kernel shared by all L latents or
    L separate kernels (gpflow SeparateIndependent + shared inducing variables)."""

    Z: np.ndarray  # [M, D]
    lengthscales: List[np.ndarray]  # per kernel: scalar () or [D]
    variances: List[float]
    q_mu: np.ndarray  # [M, L]
    q_sqrt: np.ndarray  # [L, M, M]
    mean_c: Optional[np.ndarray] = None  # [L] or None (Zero mean function)
    noise: Optional[np.ndarray] = None  # [L] Gaussian likelihood variance (experts only)
    active_dims: Optional[List[Optional[np.ndarray]]] = None  # per kernel: input columns it reads (None: all)

    def active(self, i):
        return None if self.active_dims is None or self.active_dims[i] is None else np.asarray(self.active_dims[i], dtype=np.int64)

    @property
    def M(self):
        return self.Z.shape[0]

    @property
    def L(self):
        return self.q_mu.shape[1]


@dataclass
class GPGrads:
    Z: np.ndarray
    lengthscales: List[np.ndarray]
    variances: List[np.ndarray]
    q_mu: np.ndarray
    q_sqrt: np.ndarray
    mean_c: np.ndarray
    noise: Optional[np.ndarray]


@dataclass
class ElboResult:
    elbo: float
    var_exp: float
    kl_gating: float
    kl_experts: float
    experts: Optional[List[GPGrads]] = None  # d ELBO / d (constrained parameter)
    gating: Optional[GPGrads] = None
    flat_grads: Optional[DeviceBuffer] = field(default=None, repr=False)


class Engine:
    def __init__(self, experts: List[GPBlock], gating: GPBlock, flavour="keras", max_batch=1024, device=0,
                 jitter=DEFAULT_JITTER, precision="float64"):
        """precision of the M x M x N products of predict() and of gradient-free elbo() calls: "float64" (gpflow
        default_float; FP64 DMMA), "int8" / "int8x6" (INT8 tensor cores, integer-sliced operands with exact accumulation:
        8 slices are float64-accurate, 6 slices ~1e-7), "tf32" (3xTF32 on the tcgen05 tensor cores, FP32-class accuracy).
        Calls with gradients always produce float64-accurate results: FP64 DMMA, and -- for more than 448 inducing points and
        >= 2048 rows -- the dense products on the INT8 tensor cores with eight exact byte slices (set_int8_backward(False)
        keeps everything on the FP64 pipe)."""
        self.lib = _lib.load()
        if precision not in _lib.PRECISION_IDS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISION_IDS)}")
        self.precision = precision
        self.flavour = flavour
        self.device = device
        self.K = len(experts)
        self.D = int(experts[0].Z.shape[1])
        self.F = int(experts[0].L)
        self.G = int(gating.L)
        for e in experts:
            if e.L != self.F or e.Z.shape[1] != self.D:
                raise ValueError("all experts must share input_dim and num_latent_gps")
        if gating.Z.shape[1] != self.D:
            raise ValueError("gating network input_dim differs from the experts'")
        if not ((self.G == 1 and self.K == 2) or self.G == self.K):
            raise ValueError(f"gating network with {self.G} function(s) cannot gate {self.K} experts")
        # groups: one per (GP, kernel)
        self.blocks_meta = []  # per GP: (first_group, n_groups, first_latent, L, M)
        group_M, latent_group = [], []
        for gp in list(experts) + [gating]:
            nk = len(gp.lengthscales)
            if nk not in (1, gp.L):
                raise ValueError("a GP needs 1 shared kernel or one kernel per latent")
            g0, l0 = len(group_M), len(latent_group)
            group_M += [gp.M] * nk
            latent_group += [g0 + (0 if nk == 1 else l) for l in range(gp.L)]
            self.blocks_meta.append((g0, nk, l0, gp.L, gp.M))
        self.Q, self.P, self.Pe = len(group_M), len(latent_group), self.K * self.F
        self._gm = (C.c_int32 * self.Q)(*group_M)
        self._lg = (C.c_int32 * self.P)(*latent_group)
        self.max_batch = int(max_batch)
        desc = _lib.Desc(self.D, self.F, self.K, self.G, self.Q, self.P, self.max_batch,
                         _lib.FLAVOUR_IDS[flavour], device, _lib.PRECISION_IDS[precision], jitter, self._gm, self._lg)
        h = C.c_void_p()
        _lib.check(self.lib.mogpe_create(C.byref(desc), C.byref(h)))
        self.h = h
        lo = _lib.Layout()
        _lib.check(self.lib.mogpe_param_layout(self.h, C.byref(lo)), self.h)
        self.lo = lo
        self.Mp = lo.Mp
        self.params = DeviceBuffer((lo.total,), device)
        # gradients and the 4 result scalars share one buffer so data-parallel ranks sum both in ONE all-reduce
        self.gbuf = DeviceBuffer((lo.total + 4,), device)
        self.grads_ptr = self.gbuf.ptr
        self.out_ptr = self.gbuf.ptr + lo.total * 8
        self._eps_cache = {}
        self.set_params(experts, gating)

    def close(self):
        if getattr(self, "h", None):
            self.lib.mogpe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- packing ---------------------------------------------------------------------------------
    def pack(self, experts: List[GPBlock], gating: GPBlock, inactive: float = INACTIVE_LENGTHSCALE) -> np.ndarray:
        """-> the packed buffer.  `inactive`: what the lengthscale slots outside a kernel's active_dims receive (the huge
        lengthscale for parameter values; 0 when variable ids / unconstrained values are packed: those slots are no variables)."""
        lo, Mp, D = self.lo, self.Mp, self.D
        flat = np.zeros(lo.total)
        ls = flat[lo.lengthscales:lo.lengthscales + self.Q * D].reshape(self.Q, D)
        kv = flat[lo.variance:lo.variance + self.Q]
        Z = flat[lo.Z:lo.Z + self.Q * Mp * D].reshape(self.Q, Mp, D)
        qm = flat[lo.q_mu:lo.q_mu + self.P * Mp].reshape(self.P, Mp)
        qs = flat[lo.q_sqrt:lo.q_sqrt + self.P * Mp * Mp].reshape(self.P, Mp, Mp)
        mc = flat[lo.mean_c:lo.mean_c + self.P]
        nz = flat[lo.noise:lo.noise + self.Pe]
        for gp, (g0, nk, l0, L, M) in zip(list(experts) + [gating], self.blocks_meta):
            for i in range(nk):
                ad = gp.active(i)
                if ad is None:
                    ls[g0 + i] = np.broadcast_to(np.asarray(gp.lengthscales[i], dtype=np.float64), (D,))
                else:
                    ls[g0 + i] = inactive
                    ls[g0 + i, ad] = np.broadcast_to(np.asarray(gp.lengthscales[i], dtype=np.float64), (ad.size,))
                kv[g0 + i] = float(gp.variances[i])
                Z[g0 + i, :M] = gp.Z
            qm[l0:l0 + L, :M] = np.asarray(gp.q_mu).T
            qs[l0:l0 + L, :M, :M] = np.tril(np.asarray(gp.q_sqrt))
            if gp.mean_c is not None:
                mc[l0:l0 + L] = np.broadcast_to(np.asarray(gp.mean_c, dtype=np.float64).reshape(-1), (L,))
            if gp.noise is not None:
                nz[l0:l0 + L] = np.broadcast_to(np.asarray(gp.noise, dtype=np.float64).reshape(-1), (L,))
        return flat

    def set_params(self, experts, gating):
        self._ls_shapes = [[np.shape(ls) for ls in gp.lengthscales] for gp in list(experts) + [gating]]
        self._ls_active = [[gp.active(i) for i in range(len(gp.lengthscales))] for gp in list(experts) + [gating]]
        self.params.copy_from(self.pack(experts, gating))

    def unpack_grads(self, flat: np.ndarray):
        lo, Mp, D = self.lo, self.Mp, self.D
        ls = flat[lo.lengthscales:lo.lengthscales + self.Q * D].reshape(self.Q, D)
        kv = flat[lo.variance:lo.variance + self.Q]
        Z = flat[lo.Z:lo.Z + self.Q * Mp * D].reshape(self.Q, Mp, D)
        qm = flat[lo.q_mu:lo.q_mu + self.P * Mp].reshape(self.P, Mp)
        qs = flat[lo.q_sqrt:lo.q_sqrt + self.P * Mp * Mp].reshape(self.P, Mp, Mp)
        mc = flat[lo.mean_c:lo.mean_c + self.P]
        nz = flat[lo.noise:lo.noise + self.Pe]
        out = []
        for bi, (g0, nk, l0, L, M) in enumerate(self.blocks_meta):
            gl = []
            for i in range(nk):
                shp, ad = self._ls_shapes[bi][i], self._ls_active[bi][i]
                row = ls[g0 + i] if ad is None else ls[g0 + i][ad]
                gl.append(row.sum() if shp == () else row.copy())
            out.append(GPGrads(
                Z=Z[g0:g0 + nk, :M].sum(0),
                lengthscales=gl,
                variances=[kv[g0 + i].copy() for i in range(nk)],
                q_mu=qm[l0:l0 + L, :M].T.copy(),
                q_sqrt=np.tril(qs[l0:l0 + L, :M, :M]),
                mean_c=mc[l0:l0 + L].copy(),
                noise=nz[l0:l0 + L].copy() if bi < self.K else None,
            ))
        return out[:-1], out[-1]

    # ---- eps helpers ---------------------------------------------------------------------------------
    def pack_eps_u(self, eps_u_experts, eps_u_gating, S):
        """[K][S, F, M_k] (+ gating [S, G, M_g]) -> [P, S, Mp] zero padded."""
        flat = np.zeros((self.P, S, self.Mp))
        srcs = list(eps_u_experts) if eps_u_experts is not None else [None] * self.K
        srcs.append(eps_u_gating)
        for e, (g0, nk, l0, L, M) in zip(srcs, self.blocks_meta):
            if e is not None:
                flat[l0:l0 + L, :, :M] = np.transpose(np.asarray(e, dtype=np.float64), (1, 0, 2))
        return flat

    def pack_eps_f(self, eps_f):
        """[S, N, F, K] -> [S, N, K*F] (latent order k*F + f)."""
        e = np.asarray(eps_f, dtype=np.float64)
        return np.ascontiguousarray(np.transpose(e, (0, 1, 3, 2)).reshape(e.shape[0], e.shape[1], self.Pe))

    def _dev(self, name, arr):
        """Upload a host array into a cached device buffer; pass device tensors through."""
        if arr is None:
            return None, None
        v = as_view(arr)
        if v.on_device:
            return v.ptr, v
        n = int(np.prod(v.shape))
        buf = self._eps_cache.get(name)
        if buf is None or buf.size < n:
            buf = DeviceBuffer((n,), self.device, zero=False)
            self._eps_cache[name] = buf
        _lib.check(self.lib.mogpe_memcpy_h2d(self.device, buf.ptr, v.ptr, n * 8, None))
        return buf.ptr, (buf, v)

    def fill_normal(self, buf: DeviceBuffer, seed: int, offset: int = 0, count=None, stream=None):
        _lib.check(self.lib.mogpe_fill_normal(self.h, buf.ptr, buf.size if count is None else count, seed, offset, stream), self.h)

    # ---- the hot path ----------------------------------------------------------------------------------
    def gather_rows(self, src: DeviceBuffer, idx_dev_ptr: int, n: int, cols: int, dst: DeviceBuffer, stream=None):
        """dst[i, :] = src[idx[i], :] on the device (idx: device pointer to n int64)."""
        _lib.check(self.lib.mogpe_gather_rows(self.h, src.ptr, idx_dev_ptr, int(n), int(cols), dst.ptr, stream), self.h)

    def check_numerics(self, stream=None):
        """Raise MogpeError(MOGPE_ERR_NUMERIC) if a Cholesky factorisation failed since the last check."""
        _lib.check(self.lib.mogpe_check_numerics(self.h, stream), self.h)

    def set_mc(self, eps_mc=None, num_mc=100):
        """Softmax Monte-Carlo draws for the tight bound (G > 1): eps_mc [S, num_mc, N, G] or None (library stream)."""
        ptr = None
        if eps_mc is not None:
            e = np.ascontiguousarray(eps_mc, dtype=np.float64)
            num_mc = e.shape[1]
            ptr, self._mc_keep = self._dev("eps_mc_elbo", e)
        _lib.check(self.lib.mogpe_set_mc(self.h, ptr, int(num_mc)), self.h)

    def elbo(self, X, Y, bound="further_gating", num_samples=1, scale=1.0, eps_u=None, eps_h=None, eps_f=None,
             want_grads=True, unpack=True, readback=True, stream=None) -> ElboResult:
        """eps_u [P, S, Mp], eps_h [S, N, G], eps_f [S, N, K*F] (already in library layout); any eps left None
        is drawn by the library's own Philox stream.  readback=False (device inputs only) leaves the result
        in self.gbuf and returns None without synchronising."""
        xv, yv = as_view(X), as_view(Y)
        N = int(xv.shape[0])
        if len(xv.shape) != 2 or xv.shape[1] != self.D:
            raise ValueError(f"X must be [N, {self.D}], got {xv.shape}")
        if tuple(yv.shape) != (N, self.F):
            raise ValueError(f"Y must be [{N}, {self.F}], got {tuple(yv.shape)}")
        b = _lib.BOUND_IDS[bound]
        gptr = self.grads_ptr if want_grads else None
        # all-host inputs with a synchronous read-back take the single-call host entry (one staged H2D copy); anything else --
        # device tensors, readback=False, data parallel (the results must land behind the gradients in self.gbuf so that
        # ONE all-reduce sums both) -- goes through the device entry, host arrays being uploaded first
        host_path = readback and getattr(self, "world_size", 1) == 1 and (not xv.on_device) and (not yv.on_device) and all(
            e is None or not as_view(e).on_device for e in (eps_u, eps_h, eps_f))
        if host_path:
            out = np.zeros(4)
            ev = [None if e is None else as_view(e) for e in (eps_u, eps_h, eps_f)]
            _lib.check(self.lib.mogpe_elbo_fwd_bwd_host(
                self.h, xv.ptr, yv.ptr, N, self.params.ptr, *[None if e is None else e.ptr for e in ev],
                b, num_samples, float(scale), out.ctypes.data, gptr, stream), self.h)
        else:
            keep = []
            ptrs = []
            for name, e in (("X", X), ("Y", Y), ("eps_u", eps_u), ("eps_h", eps_h), ("eps_f", eps_f)):
                p, k = self._dev(name, e)
                ptrs.append(p)
                keep.append(k)
            _lib.check(self.lib.mogpe_elbo_fwd_bwd(
                self.h, ptrs[0], ptrs[1], N, self.params.ptr, ptrs[2], ptrs[3], ptrs[4],
                b, num_samples, float(scale), self.out_ptr, gptr, stream), self.h)
            out = self.gbuf.numpy(offset=self.lo.total, count=4) if readback else None
            if readback:
                self.check_numerics(stream)
        if out is None:
            return None
        res = ElboResult(*[float(v) for v in out])
        if want_grads:
            res.flat_grads = self.gbuf
            if unpack:
                res.experts, res.gating = self.unpack_grads(self.gbuf.numpy(count=self.lo.total))
        return res

    def elbo_host_async(self, X, Y, bound, num_samples, scale, out_pinned_ptr: int, stream=None):
        """Enqueue one fwd+bwd step on host X / Y without synchronising; the 4 result doubles land in the pinned
        buffer at out_pinned_ptr once the stream reaches them."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        _lib.check(self.lib.mogpe_elbo_fwd_bwd_host_async(self.h, X.ctypes.data, Y.ctypes.data, X.shape[0], self.params.ptr,
                                                          _lib.BOUND_IDS[bound], num_samples, float(scale), self.out_ptr,
                                                          out_pinned_ptr, self.grads_ptr, stream), self.h)

    def predict(self, X, want=("f_mean", "f_var", "mix_probs", "y_mean", "y_var"), eps_mc=None, stream=None, num_mc=100):
        """-> dict of numpy arrays (f_mean/f_var [N, P], mix_probs [N, K], y_mean/y_var [N, F]).
        Softmax gating: eps_mc [R, N, K] explicit Monte-Carlo draws, or None -> num_mc draws per point from the library's
        random stream (generated on the device)."""
        xv = as_view(X)
        N = int(xv.shape[0])
        if len(xv.shape) != 2 or xv.shape[1] != self.D:
            raise ValueError(f"X must be [N, {self.D}], got {xv.shape}")
        xp, keep = self._dev("Xp", X)
        shapes = {"f_mean": (N, self.P), "f_var": (N, self.P), "mix_probs": (N, self.K),
                  "y_mean": (N, self.F), "y_var": (N, self.F)}
        bufs = {k: DeviceBuffer(shapes[k], self.device, zero=False) for k in want}
        num_mc = int(num_mc) if self.G > 1 else 0
        mp = None
        if eps_mc is not None:
            e = np.ascontiguousarray(eps_mc, dtype=np.float64)
            if e.shape[1:] != (N, self.K):
                raise ValueError(f"eps_mc must be [R, {N}, {self.K}]")
            num_mc = e.shape[0]
            mp, keep2 = self._dev("eps_mc", e)
        args = [bufs[k].ptr if k in bufs else None for k in ("f_mean", "f_var", "mix_probs", "y_mean", "y_var")]
        _lib.check(self.lib.mogpe_predict(self.h, xp, N, self.params.ptr, *args, mp, num_mc, stream), self.h)
        self.check_numerics(stream)
        return {k: b.numpy() for k, b in bufs.items()}

    # ---- random stream / data parallel / measurement hooks ------------------------------------------------
    def set_seed(self, seed: int):
        _lib.check(self.lib.mogpe_set_seed(self.h, int(seed)), self.h)

    def init_data_parallel(self, rank: int, world_size: int, unique_id: bytes, overlap: bool = True):
        """Join an NCCL communicator (id from Engine.nccl_unique_id() on rank 0, distributed by the caller) and
        weight the redundantly computed KL terms by 1/world_size so that the all-reduced sums are exact.
        Random draws the caller does not supply: use the SAME seed on every rank -- the inducing-point draws (eps_u) then
        come from a shared stream (identical on all ranks, as one process would draw them) and the per-point draws (eps_h,
        eps_f, Softmax Monte-Carlo) from a per-rank stream, so the shards are independent.
        overlap: elbo() itself all-reduces [gradients | results] (most of it underneath the Cholesky backward) and
        allreduce(self.gbuf) becomes a no-op; False keeps the reduction a separate call after elbo()."""
        _lib.check(self.lib.mogpe_comm_init(self.h, unique_id, rank, world_size), self.h)
        _lib.check(self.lib.mogpe_set_kl_weight(self.h, 1.0 / world_size), self.h)
        _lib.check(self.lib.mogpe_set_dp_overlap(self.h, 1 if overlap else 0), self.h)
        self.dp_overlap = bool(overlap)
        self.world_size = world_size

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _lib.check(_lib.load().mogpe_nccl_unique_id(buf))
        return buf.raw

    def allreduce(self, buf: DeviceBuffer, count=None, stream=None):
        if getattr(self, "dp_overlap", False) and buf is self.gbuf:
            return  # elbo() already returned the globally reduced [gradients | results]
        _lib.check(self.lib.mogpe_allreduce_sum(self.h, buf.ptr, buf.size if count is None else count, stream), self.h)

    def profile(self, on=True):
        _lib.check(self.lib.mogpe_profile_enable(self.h, 1 if on else 0), self.h)

    def profile_read(self):
        ms, fl = C.c_double(), C.c_double()
        ng, na = C.c_int64(), C.c_int64()
        _lib.check(self.lib.mogpe_profile_read(self.h, C.byref(ms), C.byref(fl), C.byref(ng), C.byref(na)), self.h)
        return {"gemm_ms": ms.value, "gemm_flops": fl.value, "gemm_launches": ng.value, "all_launches": na.value}

    def profile_read_streaming(self):
        """-> {kernel: (ms, algorithmic bytes)} for the streaming kernels timed in profiling mode."""
        ms, by = (C.c_double * 4)(), (C.c_double * 4)()
        _lib.check(self.lib.mogpe_profile_read_streaming(self.h, ms, by), self.h)
        return {k: (ms[i], by[i]) for i, k in enumerate(("kuf", "lik", "build_abar", "kuf_grad"))}

    # ---- device-side optimiser -------------------------------------------------------------------
    def setup_optimizer(self, var_rep: np.ndarray, noise_lower: float = 1e-6):
        """var_rep [total] int32: representative packed index of each entry's variable (-1: not a variable)."""
        rep = np.ascontiguousarray(var_rep, dtype=np.int32)
        if rep.shape != (self.lo.total,):
            raise ValueError(f"var_rep must have {self.lo.total} entries")
        _lib.check(self.lib.mogpe_opt_setup(self.h, rep.ctypes.data_as(C.POINTER(C.c_int32)), float(noise_lower)), self.h)
        if getattr(self, "u", None) is None:
            self.u = DeviceBuffer((self.lo.total,), self.device)

    def get_optimizer_state(self):
        """-> (m, v): Adam's moment estimates in the packed layout (host copies)."""
        m, v = np.empty(self.lo.total), np.empty(self.lo.total)
        _lib.check(self.lib.mogpe_opt_state(self.h, m.ctypes.data, v.ctypes.data, 0), self.h)
        return m, v

    def set_optimizer_state(self, m: np.ndarray, v: np.ndarray):
        m, v = np.ascontiguousarray(m, dtype=np.float64), np.ascontiguousarray(v, dtype=np.float64)
        if m.shape != (self.lo.total,) or v.shape != (self.lo.total,):
            raise ValueError("optimizer state does not match this model's packed layout")
        _lib.check(self.lib.mogpe_opt_state(self.h, m.ctypes.data, v.ctypes.data, 1), self.h)

    def set_unconstrained(self, u_flat: np.ndarray, stream=None):
        """Upload the unconstrained variables and refresh params = T(u) on the device."""
        self.u.copy_from(u_flat)
        _lib.check(self.lib.mogpe_opt_transform(self.h, self.u.ptr, self.params.ptr, stream), self.h)

    def adam_step(self, t: int, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, stream=None):
        """One Adam step on loss = -ELBO from the gradients of the last elbo() call (all on the device)."""
        _lib.check(self.lib.mogpe_opt_adam_step(self.h, self.u.ptr, self.params.ptr, self.grads_ptr, float(lr), float(beta_1),
                                                float(beta_2), float(epsilon), int(t), stream), self.h)

    def train_step(self, X, Y, bound, num_samples, scale, t, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7,
                   out_pinned_ptr=None, use_graph=True):
        """One fused training step (ELBO fwd+bwd -> Adam -> params) on the library's stream, replayed as a CUDA graph from
        its third occurrence (use_graph: False / 0 never, True / 1 when the step is launch-bound, 2 always).
        X / Y: host numpy arrays (staged through pinned memory) or device TensorViews."""
        if isinstance(X, np.ndarray):
            X = np.ascontiguousarray(X, dtype=np.float64)
            Y = np.ascontiguousarray(Y, dtype=np.float64)
            xp, yp, on_host, n = X.ctypes.data, Y.ctypes.data, 1, X.shape[0]
        else:
            xv, yv = as_view(X), as_view(Y)
            xp, yp, on_host, n = xv.ptr, yv.ptr, 0, int(xv.shape[0])
        _lib.check(self.lib.mogpe_train_step(self.h, xp, yp, on_host, n, self.params.ptr, _lib.BOUND_IDS[bound], num_samples,
                                             float(scale), self.out_ptr, out_pinned_ptr, self.grads_ptr, self.u.ptr, float(lr),
                                             float(beta_1), float(beta_2), float(epsilon), int(t), int(use_graph)), self.h)

    def train_step_rows(self, X_all, Y_all, rows, bound, num_samples, scale, t, lr=0.001, beta_1=0.9, beta_2=0.999,
                        epsilon=1e-7, out_pinned_ptr=None, use_graph=True):
        """train_step on the rows `rows` (int64 index array) of the HOST data set X_all / Y_all (C-contiguous float64): the
        library gathers them straight into its pinned staging slot."""
        _lib.check(self.lib.mogpe_train_step_rows(self.h, X_all.ctypes.data, Y_all.ctypes.data, rows.ctypes.data, int(rows.shape[0]),
                                                  self.params.ptr, _lib.BOUND_IDS[bound], num_samples, float(scale), self.out_ptr,
                                                  out_pinned_ptr, self.grads_ptr, self.u.ptr, float(lr), float(beta_1),
                                                  float(beta_2), float(epsilon), int(t), int(use_graph)), self.h)

    def set_int8_backward(self, on: bool):
        """Lbar / Sbar of the backward on the INT8 tensor cores (integer-sliced, exact accumulation) when the shape qualifies
        (default on); off keeps them on the FP64 DMMA pipe."""
        _lib.check(self.lib.mogpe_set_int8_backward(self.h, 1 if on else 0), self.h)

    def set_int8_step(self, on: bool):
        """The whole gradient step (all six dense products) on the INT8 tensor cores when the shape qualifies (default on;
        requires set_int8_backward(True)); off keeps the forward / Abar / W products on the FP64 DMMA pipe."""
        _lib.check(self.lib.mogpe_set_int8_step(self.h, 1 if on else 0), self.h)

    def profile_read_int8(self):
        ms, fl, n, ops = C.c_double(), C.c_double(), C.c_int64(), C.c_double()
        _lib.check(self.lib.mogpe_profile_read_int8(self.h, C.byref(ms), C.byref(fl), C.byref(n), C.byref(ops)), self.h)
        return {"ms": ms.value, "flops": fl.value, "launches": n.value, "tensor_ops": ops.value}

    def set_int8_grad_slices(self, slices: int):
        """Byte slices of the gradient-only INT8 products (S LTA, W, Lbar, Sbar): 7 (default), 8 or 6; capped by the forward
        products' count (set_int8_forward_slices)."""
        _lib.check(self.lib.mogpe_set_int8_grad_slices(self.h, int(slices)), self.h)

    def set_int8_forward_slices(self, slices: int):
        """Byte slices of the two forward INT8 products (A = L^-1 Kuf, S^T A): 8 (default) or 7 -- both float64-accurate at the
        1e-8 / 1e-6 tolerances -- or 6: fast training mode, every product of the step loads six slices (42 bits; ELBO ~1e-9,
        gradients ~1e-6 of the float64 step: the "1e-4 class" of BASELINE.json)."""
        _lib.check(self.lib.mogpe_set_int8_forward_slices(self.h, int(slices)), self.h)

    def graph_stats(self):
        """-> (captured training-step graphs, capture disabled after a failure)."""
        n, d = C.c_int32(), C.c_int32()
        _lib.check(self.lib.mogpe_graph_stats(self.h, C.byref(n), C.byref(d)), self.h)
        return n.value, bool(d.value)

    def predict_f_inducing_samples(self, X, num_samples, eps_u=None, stream=None):
        """-> f_mean [S, N, P], f_var [N, P] (mogpe/keras/gps.py:14-50); eps_u [P, S, Mp] or None (library stream)."""
        xv = as_view(X)
        N = int(xv.shape[0])
        if len(xv.shape) != 2 or xv.shape[1] != self.D:
            raise ValueError(f"X must be [N, {self.D}], got {xv.shape}")
        xp, keep = self._dev("Xp", X)
        ep, keep2 = self._dev("eps_u", eps_u)
        fm = DeviceBuffer((num_samples, N, self.P), self.device, zero=False)
        fv = DeviceBuffer((N, self.P), self.device, zero=False)
        _lib.check(self.lib.mogpe_predict_f_inducing_samples(self.h, xp, N, self.params.ptr, ep, num_samples, fm.ptr, fv.ptr, stream), self.h)
        return fm.numpy(), fv.numpy()

    def prior_kl(self, stream=None):
        """-> KL of every latent GP [P]."""
        kl = DeviceBuffer((self.P,), self.device)
        _lib.check(self.lib.mogpe_prior_kl(self.h, self.params.ptr, kl.ptr, stream), self.h)
        return kl.numpy()

    def debug(self, name):
        n = C.c_int64()
        _lib.check(self.lib.mogpe_debug_copy(self.h, name.encode(), None, 0, C.byref(n)), self.h)
        out = np.empty(n.value)
        _lib.check(self.lib.mogpe_debug_copy(self.h, name.encode(), out.ctypes.data, n.value, C.byref(n)), self.h)
        return out
