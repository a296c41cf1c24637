#!/usr/bin/env python3
"""Benchmark of the mogpe hot path: minibatch ELBO forward+backward of MixtureOfSVGPExperts.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C4] [--impl ours|reference] [--scaling weak|strong]

One "step" = one pass of the hot path over one synthetic minibatch: draw the Monte-Carlo normals on the
device, ELBO forward + full backward (gradients of every kernel / inducing / variational parameter), and
for N > 1 one NCCL all-reduce of the packed gradient buffer.  Prints ONE JSON line (rank 0).

Workloads (per-GPU batch fixed -> weak scaling):
  C1 mcycle-like  D=1 F=1 K=2 G=1 M=32   batch 16
  C2 artificial   D=1 F=1 K=3 G=3 M=128  batch 8192
  C3 quadcopter   D=2 F=2 K=2 G=1 M=256  batch 16384
  C4 (default)    D=4 F=1 K=8 G=8 M=512  batch 8192 per GPU (65536 global on 8 GPUs)  <- BASELINE.json metric
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # `batch` = rows per GPU under weak scaling; `strong_global` = the fixed global batch of --scaling strong
    "C1": dict(D=1, F=1, K=2, M=32, batch=16, num_data=93, sep=False, strong_global=16),
    "C2": dict(D=1, F=1, K=3, M=128, batch=8192, num_data=100_000, sep=False, strong_global=8192),
    "C3": dict(D=2, F=2, K=2, M=256, batch=16384, num_data=1_000_000, sep=True, strong_global=16384),
    "C4": dict(D=4, F=1, K=8, M=512, batch=8192, num_data=10_000_000, sep=False, strong_global=65536),
}
BOUND, FLAVOUR, S = "further_gating", "keras", 1
FP64_TENSOR_PEAK_FALLBACK = 37.0  # DMMA issue peak measured on this pool in round 1 (profiles/fp64_probe_r01.log)


def fp64_tensor_peak(device):
    """FP64 tensor (DMMA) issue peak measured NOW on this GPU by the library's probe kernel (MEASURED_PEAKS.json holds only
    HBM and bf16 figures and the profiling recipe states no FP64 fallback)."""
    import ctypes as C

    from mogpe_b200 import _lib

    v = C.c_double()
    _lib.check(_lib.load().mogpe_probe_fp64_tensor_peak(device, C.byref(v)))
    return float(v.value)


def int8_tensor_peak(device):
    """INT8 tensor-pipe issue peak (TOP/s) measured NOW by the library's probe: back-to-back tcgen05.mma.kind::i8 128x256x32
    from shared-memory operands, one issuing warp per SM."""
    import ctypes as C

    from mogpe_b200 import _lib

    v = C.c_double()
    _lib.check(_lib.load().mogpe_probe_int8_tensor_peak(device, C.byref(v)))
    return float(v.value)


def measured_traffic(kernel_key):
    """dram read+write bytes per launch of the dominant kernel from the committed `ncu --set full` summary of this round
    (profiles/traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); None when no capture is committed."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        e = t[kernel_key]
        return float(e["dram_bytes_per_launch"]), f"{e['source']} ({e.get('what', kernel_key)}); algorithmic {e.get('algorithmic_bytes')}"
    except Exception as exc:  # loud, not silent: the line says the figure is missing instead of repeating a stale constant
        return None, f"NO COMMITTED ncu CAPTURE for '{kernel_key}' in profiles/traffic.json ({type(exc).__name__})"


def _hbm_peak():
    """Measured copy bandwidth of this pool's B200 (driver-written MEASURED_PEAKS.json), else the profiling recipe's
    fallback."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6543.7


HBM_PEAK_GBS = _hbm_peak()


def logu(rng, lo, hi, size=None):
    return np.exp(rng.uniform(np.log(lo), np.log(hi), size))


def synth_model(wl, seed=3):
    """Model parameters in the oracle's spec format: l, var ~ LogU(0.3, 3), noise ~ LogU(0.3, 1)
    (kept where the reference's probability-space ELBO is finite), q_mu ~ N(0,1), q_sqrt = tril(0.1 N) + I,
    Z = M data-like rows."""
    rng = np.random.default_rng(seed)
    D, F, K, M = wl["D"], wl["F"], wl["K"], wl["M"]
    G = 1 if K == 2 else K

    def gp(L, nk, expert):
        d = {
            "Z": rng.standard_normal((M, D)),
            "kernels": [{"lengthscales": logu(rng, 0.3, 3.0, D), "variance": float(logu(rng, 0.3, 3.0))} for _ in range(nk)],
            "q_mu": rng.standard_normal((M, L)),
            "q_sqrt": np.tril(0.1 * rng.standard_normal((L, M, M))) + np.eye(M)[None],
            "mean_c": 0.1 * rng.standard_normal(L) if expert else None,
        }
        if expert:
            d["noise_variance"] = logu(rng, 0.3, 1.0, L)
        return d

    experts = [gp(F, F if wl["sep"] else 1, True) for _ in range(K)]
    gating = gp(G, G, False)  # G separate gating kernels + shared Z ("3 RBF kernels, shared Z")
    return {"flavour": FLAVOUR, "bound": BOUND, "num_samples": S, "num_data": wl["num_data"],
            "experts": experts, "gating": gating}


def synth_batch(wl, n, seed):
    rng = np.random.default_rng(seed)
    D, F, K = wl["D"], wl["F"], wl["K"]
    X = rng.standard_normal((n, D))
    regime = rng.integers(0, K, n)
    Y = np.stack([np.sin((1 + f) * X.sum(-1) + regime) + 0.1 * (1 + regime) * rng.standard_normal(n) for f in range(F)], -1)
    return np.ascontiguousarray(X), np.ascontiguousarray(Y)


def blocks_from_spec(spec):
    from mogpe_b200.engine import GPBlock

    def blk(g, expert):
        L = np.asarray(g["q_mu"]).shape[1]
        return GPBlock(Z=np.asarray(g["Z"], float), lengthscales=[np.asarray(k["lengthscales"], float) for k in g["kernels"]],
                       variances=[float(k["variance"]) for k in g["kernels"]], q_mu=np.asarray(g["q_mu"], float),
                       q_sqrt=np.asarray(g["q_sqrt"], float),
                       mean_c=None if g.get("mean_c") is None else np.asarray(g["mean_c"], float),
                       noise=np.broadcast_to(np.asarray(g["noise_variance"], float).reshape(-1), (L,)).copy() if expert else None)

    return [blk(g, True) for g in spec["experts"]], blk(spec["gating"], False)


def build_public_model(spec, wl, device=0):
    """The same model through the drop-in API classes (what a user of the reference constructs)."""
    from mogpe_b200 import gpflow_lite as gl
    from mogpe_b200.keras import MixtureOfSVGPExperts, SVGPExpert, SVGPGatingNetwork

    def kern(g, L):
        ks = [gl.RBF(variance=k["variance"], lengthscales=np.asarray(k["lengthscales"], float)) for k in g["kernels"]]
        return ks[0] if len(ks) == 1 and L == 1 else (gl.SharedIndependent(ks[0], L) if len(ks) == 1 else gl.SeparateIndependent(ks))

    experts = []
    for g in spec["experts"]:
        L = np.asarray(g["q_mu"]).shape[1]
        k = kern(g, L)
        k = k.kernel if isinstance(k, gl.SharedIndependent) else k
        experts.append(SVGPExpert(k, gl.Gaussian(np.asarray(g["noise_variance"], float).reshape(1, -1)), gl.InducingPoints(g["Z"]),
                                  gl.Constant(np.asarray(g["mean_c"], float)), num_latent_gps=L, q_mu=g["q_mu"], q_sqrt=g["q_sqrt"]))
    g = spec["gating"]
    G = np.asarray(g["q_mu"]).shape[1]
    gating = SVGPGatingNetwork(kern(g, G), gl.InducingPoints(g["Z"]), q_mu=g["q_mu"], q_sqrt=g["q_sqrt"])
    return MixtureOfSVGPExperts(experts, gating, num_samples=S, num_data=wl["num_data"], bound=BOUND, device=device)


def step_flops(wl, B):
    """Algorithmic FLOPs of one fwd+bwd step (BASELINE.md work model; FMA = 2)."""
    F, K, M = wl["F"], wl["K"], wl["M"]
    G = 1 if K == 2 else K
    Pe, Pg = K * F, G
    return Pe * (4.0 * M * M * B + 8.0 * M**3 / 3.0) + Pg * (8.0 * M * M * B + 8.0 * M**3 / 3.0)


# ---- the CPU arm: the restated reference (oracle/ref_torch.py) ------------------------------------------
def run_cpu_reference(wl, spec, steps, warmup, rows):
    import torch

    from oracle import ref_torch as rt

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = rt.MixtureOfSVGPExperts(spec)
    X, Y = synth_batch(wl, rows, seed=1000)
    rng = np.random.default_rng(11)
    K, F, M = wl["K"], wl["F"], wl["M"]
    G = 1 if K == 2 else K
    eps = {"u_experts": [rng.standard_normal((S, F, M)) for _ in range(K)], "h": rng.standard_normal((S, rows, G))}
    for _ in range(warmup):
        model.elbo_and_grads(X, Y, eps)
    t0 = time.perf_counter()
    for _ in range(steps):
        model.elbo_and_grads(X, Y, eps)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return {"value": rows / dt, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} fwd+bwd steps of {rows} rows of workload (same model, torch float64 autograd "
                      f"following the reference's per-expert op sequence), {dt * 1e3:.1f} ms/step"}, dt


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def run_predict_modes(device, n_points=262144, M=1024, fp64_peak=None):
    """The other data path of the north star (BASELINE config 5: predict_y + predict_mixing_probs, K = 8, G = 8, M = 1024) on a
    bounded sample of test points, once per arithmetic mode of the dense products: FP64 DMMA, INT8 tensor cores with
    integer-sliced operands (8 slices: FP64-accurate; 6 slices: ~1e-7), 3xTF32 tensor cores.  Points are generated on the device; CUDA-event time."""
    import torch

    from mogpe_b200 import _lib
    from mogpe_b200.device import DeviceBuffer
    from mogpe_b200.engine import Engine

    wl = dict(D=4, F=1, K=8, M=M, batch=65536, num_data=n_points, sep=False)
    spec = synth_model(wl, seed=4)
    experts, gating = blocks_from_spec(spec)
    lib = _lib.load()
    out = {"workload": f"C5 sample: predict_y + predict_mixing_probs, D=4 K=8 G=8 M={M}, {n_points} test points in 65536-point "
                       "chunks, 100 Softmax Monte-Carlo draws per point from the library stream", "unit": "points/s"}
    ref = None
    for precision in ("float64", "int8", "int8x6", "tf32"):
        eng = Engine(experts, gating, flavour="keras", max_batch=65536, precision=precision, device=device)
        X = DeviceBuffer((n_points, 4), device, zero=False)
        eng.fill_normal(X, seed=4)
        ym, yv, mix = (DeviceBuffer(s, device, zero=False) for s in ((n_points, 1), (n_points, 1), (n_points, 8)))

        def run():
            _lib.check(lib.mogpe_predict(eng.h, X.ptr, n_points, eng.params.ptr, None, None, mix.ptr, ym.ptr, yv.ptr,
                                         None, 100, None), eng.h)

        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        eng.set_seed(0)  # same Monte-Carlo draws for the accuracy comparison below
        run()
        got = np.concatenate([ym.numpy(), yv.numpy()], 1)
        if ref is None:
            ref = got
        flops = 16 * 2.0 * M * M * n_points  # (P_e + P_g) * 2 M^2 per point
        out[precision] = {"value": n_points / (ms * 1e-3), "ms": ms, "algorithmic_tflops": flops / ms * 1e-9,
                          "frac_of_fp64_tensor_peak": flops / ms * 1e-9 / (fp64_peak or FP64_TENSOR_PEAK_FALLBACK),
                          "max_rel_diff_vs_float64": float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))}
        eng.close()
        del X, ym, yv, mix
    return out


def dp_parity_check(eng, wl, B, rank, world, pool, scale, local_rank):
    """Numerical check of the data-parallel path at the bench shape (run by every `--gpus N > 1` invocation, so the
    overlapped NCCL reduction is verified on the 8-GPU box): one extra step with FIXED draws on all ranks, against the same
    global minibatch evaluated on rank 0 alone, shard by shard, with the in-library reduction switched off and the partial
    [gradients | results] buffers summed on the host.  -> {"elbo_rel", "grad_max_rel"} on rank 0, None elsewhere."""
    import torch.distributed as dist

    from mogpe_b200 import _lib

    K, F, M = wl["K"], wl["F"], wl["M"]
    G = 1 if K == 2 else K
    r_u, r_h = np.random.default_rng(77), np.random.default_rng(78)
    eps_u = eng.pack_eps_u([r_u.standard_normal((S, F, M)) for _ in range(K)], None, S)
    eps_h = r_h.standard_normal((S, B * world, G))
    shard = lambda r: synth_batch(wl, B, seed=100 + r * pool)
    sl = lambda r: np.ascontiguousarray(eps_h[:, r * B:(r + 1) * B])
    X, Y = shard(rank)
    eng.elbo(X, Y, BOUND, S, scale, eps_u=eps_u, eps_h=sl(rank), want_grads=True, unpack=False, readback=False)
    if world > 1:
        eng.allreduce(eng.gbuf)
    got = eng.gbuf.numpy()
    out = None
    if rank == 0:
        _lib.check(eng.lib.mogpe_set_dp_overlap(eng.h, 0), eng.h)  # no collective below: ranks 1.. are idle
        acc = np.zeros_like(got)
        for r in range(world):
            Xr, Yr = shard(r)
            eng.elbo(Xr, Yr, BOUND, S, scale, eps_u=eps_u, eps_h=sl(r), want_grads=True, unpack=False, readback=False)
            acc += eng.gbuf.numpy()
        lo = eng.lo
        bounds = [lo.lengthscales, lo.variance, lo.Z, lo.q_mu, lo.q_sqrt, lo.mean_c, lo.noise, lo.total]
        gmax = 0.0
        for a, b in zip(bounds[:-1], bounds[1:]):
            gmax = max(gmax, float(np.max(np.abs(got[a:b] - acc[a:b])) / max(1e-300, np.max(np.abs(acc[a:b])))))
        out = {"elbo_rel": float(abs(got[-4] - acc[-4]) / abs(acc[-4])), "grad_max_rel": gmax,
               "elbo_dp": float(got[-4]), "elbo_single": float(acc[-4]),
               "how": f"fixed draws; {world}-rank overlapped NCCL step vs the same {B * world}-row minibatch on rank 0 alone "
                      "(shard by shard, no collective, host sum); grad_max_rel = worst block-wise max|diff| / max|grad|"}
        _lib.check(eng.lib.mogpe_set_dp_overlap(eng.h, 1 if getattr(eng, "dp_overlap", False) else 0), eng.h)
    if world > 1:
        dist.barrier()
    return out


def _arm_watchdog():
    """A multi-rank run whose ranks disagree on a collective hangs silently; give up loudly instead (default 30 min,
    MOGPE_BENCH_TIMEOUT seconds; 0 disables).  A default run takes a few minutes."""
    import threading

    limit = float(os.environ.get("MOGPE_BENCH_TIMEOUT", "1800"))
    if limit <= 0:
        return

    def expire():
        sys.stderr.write(f"bench.py: no result after {limit:.0f} s (rank {os.environ.get('RANK', '0')}): giving up\n")
        sys.stderr.flush()
        os._exit(3)

    t = threading.Timer(limit, expire)
    t.daemon = True
    t.start()


def main():
    _arm_watchdog()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="C4", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: fixed rows per GPU (default, BASELINE config 4 at 8 GPUs); strong: fixed global batch "
                         "(SURVEY 8d: 65536 rows for C4) split over the GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-predict", action="store_true")
    ap.add_argument("--fast", action="store_true",
                    help="fast training mode (north_star's 1e-4 class): six byte slices in every INT8 product; NOT the headline")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_arm = world if args.impl == "ours" else max(1, args.gpus)
    steps, warmup = args.steps, max(args.warmup, 3)
    K = wl["K"]
    G = 1 if K == 2 else K
    if args.scaling == "strong":
        if wl["strong_global"] % n_arm:
            raise SystemExit(f"--scaling strong: global batch {wl['strong_global']} is not divisible by {n_arm} GPUs")
        B = wl["strong_global"] // n_arm
    else:
        B = wl["batch"]
    cfg = {"workload": f"{args.workload}: D={wl['D']} F={wl['F']} K={K} G={G} M={wl['M']} S={S} "
                       f"bound={BOUND} flavour={FLAVOUR}, batch {B}/GPU (global {B * n_arm}), "
                       f"num_data={wl['num_data']}",
           "batch_per_gpu": B, "global_batch": B * n_arm, "M": wl["M"], "K": K, "D": wl["D"],
           "l2": "per-step working set (A, Kuf, W, ... = 4*Q*M*B*8 B) is far larger than the 126 MB L2 for C2-C4; "
                 "minibatches cycle through a 4-batch pool",
           "parallelism": f"dp{n_arm}: minibatch rows sharded, params replicated, one NCCL all-reduce of grads",
           "arithmetic": "float64 results (1e-8 / 1e-6 parity); for M >= 512 the dense products run as exact integer-sliced "
                         "INT8 tensor-core products (8 byte slices), FP64 DMMA otherwise"}
    spec = synth_model(wl)

    if args.impl == "reference":
        if rank != 0:
            return
        # the restated reference on the host cores, full per-step batch of the workload (bounded at 8192 rows: the CPU step
        # is row-linear apart from the M^3 Cholesky terms, and 65536 rows would need > 20 GB of autograd intermediates)
        rows = min(B * n_arm, 8192)
        cb, dt = run_cpu_reference(wl, spec, max(1, steps), warmup, rows)
        line = {"impl": "reference", "metric": "elbo_fwd_bwd_points_per_sec", "value": cb["value"], "unit": "points/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "restated reference (torch-CPU float64, oracle/ref_torch.py): TensorFlow/GPflow/TFP are not "
                        "installable here, so the reference's own code cannot run"}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist

    from mogpe_b200.device import DeviceBuffer
    from mogpe_b200.engine import Engine

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    experts, gating = blocks_from_spec(spec)
    eng = Engine(experts, gating, flavour=FLAVOUR, max_batch=B, device=local_rank)
    if args.fast:
        eng.set_int8_forward_slices(6)
        cfg["arithmetic"] = "fast mode: six byte slices (42 bits) in every INT8 product -- the 1e-4 class, not float64 parity"
    eng.set_seed(1234)  # one seed for the job: inducing draws are shared between the ranks, per-point draws are per rank
    if world > 1:
        ids = [Engine.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.init_data_parallel(rank, world, ids[0])
    scale = wl["num_data"] / float(B * world)
    pool = 4
    host_batches = [synth_batch(wl, B, seed=100 + rank * pool + i) for i in range(pool)]
    dev_batches = [(DeviceBuffer.from_numpy(x, local_rank), DeviceBuffer.from_numpy(y, local_rank)) for x, y in host_batches]

    def step(i):
        x, y = dev_batches[i % pool]
        eng.elbo(x, y, BOUND, S, scale, want_grads=True, unpack=False, readback=False)
        if world > 1:
            eng.allreduce(eng.gbuf)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None  # samples cover warm-up + timed region (all under load)
    t_load = time.perf_counter()
    for i in range(warmup):
        step(i)
    barrier()
    # keep the GPU under the same load until nvidia-smi has had time to report (its first sample takes ~0.3 s)
    # (every step contains a collective when world > 1: the NUMBER of these extra steps must be the same on all ranks, so rank
    # 0's clock decides and the others follow -- a per-rank time test lets one rank leave the loop a step early and deadlocks
    # the rest in their all-reduce)
    i = warmup
    while True:
        more = [time.perf_counter() - t_load < 0.6]
        if world > 1:
            dist.broadcast_object_list(more, src=0)
        if not more[0]:
            break
        step(i)
        i += 1
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.profile_read()  # zero the launch counters
    barrier()
    e0.record()
    for i in range(steps):
        step(warmup + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches_timed = eng.profile_read()["all_launches"]  # counters run even with event timing off
    # second pass over the same K steps with every dense-product launch bracketed by CUDA events (roofline numbers).
    # The events serialise the side-stream overlap, so this pass is a little slower than the timed one; `value`
    # comes from the clean pass above.
    eng.profile(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(steps):
        step(warmup + i)
    p1.record()
    barrier()
    ms_prof = p0.elapsed_time(p1)
    prof = eng.profile_read()
    prof8 = eng.profile_read_int8()
    stream_prof = eng.profile_read_streaming()
    eng.profile(False)
    clocks = sampler.stop() if sampler else None
    elbo_val = eng.gbuf.numpy(offset=eng.lo.total, count=4)
    dp_parity = dp_parity_check(eng, wl, B, rank, world, pool, scale, local_rank) if world > 1 else None
    fp64_peak = fp64_tensor_peak(local_rank) if rank == 0 else None
    i8_peak = int8_tensor_peak(local_rank) if rank == 0 else None

    # End to end through the public API a user of the reference calls: MixtureOfSVGPExperts.fit (Keras flavour).
    # Every step copies its minibatch host -> device (staged through pinned memory inside the library), runs ELBO forward +
    # backward, [all-reduces the gradients], applies the device-side Adam step, and reads the 4 result scalars back
    # (device -> host, asynchronously into pinned memory; fit() synchronises on them at the end of the epoch).
    del eng, dev_batches
    model = build_public_model(spec, wl, device=local_rank)
    model.set_seed(4321)
    if args.fast:
        model.set_fast_training(True)
    if world > 1:
        ids2 = [Engine.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids2, src=0)
        model.enable_data_parallel(rank, world, ids2[0])
    from mogpe_b200._model_core import Adam

    model.compile(optimizer=Adam(learning_rate=1e-3))
    Xh = np.concatenate([hb[0] for hb in host_batches] * ((steps + 3) // pool + 1))[: (steps + 3) * B]
    Yh = np.concatenate([hb[1] for hb in host_batches] * ((steps + 3) // pool + 1))[: (steps + 3) * B]
    model.fit(Xh[: 3 * B], Yh[: 3 * B], epochs=1, batch_size=B, shuffle=False)  # warm-up (3 steps)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    hist = model.fit(Xh[3 * B:], Yh[3 * B:], epochs=1, batch_size=B, shuffle=False)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pts = float(B * world * steps)
    value = pts / (ms * 1e-3)
    gemm_tflops = prof["gemm_flops"] / (prof["gemm_ms"] * 1e-3) * 1e-12 if prof["gemm_ms"] > 0 else None
    i8_on = prof8["ms"] > 0
    # INT8-sliced products: every FP64-equivalent multiply-add is 36 signed-byte slice-pair multiply-adds (8 x 8 slices, the
    # pairs with i + j < 8), so the tensor pipe executes 36 operations per algorithmic FLOP
    i8_eq_tflops = prof8["flops"] / (prof8["ms"] * 1e-3) * 1e-12 if i8_on else None
    i8_tops = prof8["tensor_ops"] / (prof8["ms"] * 1e-3) * 1e-12 if i8_on else None
    dmma = {"kernel": "mogpe::gemm_dmma_kernel (FP64 DMMA: the M^3 products of the Cholesky backward; every dense product when "
                      "the INT8 step does not apply)",
            "achieved": gemm_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": None if gemm_tflops is None else gemm_tflops / fp64_peak,
            "ms_per_step": prof["gemm_ms"] / steps, "launches_per_step": prof["gemm_launches"] / steps - (prof8["launches"] / steps if i8_on else 0)}
    if i8_on:
        traffic, traffic_note = measured_traffic("gemm_i8_A")
        roof = {
            "bound": "tensor", "kernel": "mogpe::oz::gemm_i8_sliced_kernel<NS, EW> (tcgen05.mma kind::i8 + TMA + TMEM; the six dense "
                                         "products of a step: A = L^-1 Kuf, S^T A, S S^T A, W = L^-T Abar, Lbar, Sbar)",
            "achieved": i8_tops, "peak": i8_peak, "unit": "TOP/s", "frac": i8_tops / i8_peak,
            "peak_source": "INT8 tensor-pipe issue peak measured in THIS run by mogpe_probe_int8_tensor_peak (back-to-back "
                           "tcgen05.mma.kind::i8 128x256x32 from shared-memory operands on every SM; nominal dense INT8: 4500); "
                           "MEASURED_PEAKS.json holds only HBM and bf16 figures",
            "achieved_note": "INT8 slice-pair operations (36 per algorithmic FP64 FLOP for the two forward products with eight "
                             "slices, 28 for the four gradient-only products with seven; triangular operands / results counted "
                             "as triangular) / device time of the launches (CUDA events around every launch, profiled pass)",
            "fp64_equivalent_tflops": i8_eq_tflops,
            "fp64_equivalent_vs_fp64_tensor_peak": i8_eq_tflops / fp64_peak,
            "kernel_ms_per_step": prof8["ms"] / steps, "kernel_share_of_step": prof8["ms"] / ms_prof,
            "kernel_launches_per_step": prof8["launches"] / steps,
            "profiled_pass_ms_per_step": ms_prof / steps,
            "step_model_tflops": step_flops(wl, B) * steps / (ms * 1e-3) * 1e-12,
            "step_model_vs_fp64_tensor_peak": step_flops(wl, B) * steps / (ms * 1e-3) * 1e-12 / fp64_peak,
            "what_bounds_it": "per 64-wide k-block a CTA ingests 96 KB of operand planes by TMA and the MMAs read 208 KB from "
                              "shared memory for 2438 tensor cycles: the main loop sits at the shared-memory bandwidth (~60 % of the "
                              "issue peak); at K <= 512 the period of a tile is TMEM drain + max(main loop, rest of the epilogue) "
                              "and the epilogue (FP64 re-slicing, statistics, plane stores) is the longer one (DESIGN.md 4d)",
            "traffic": traffic, "traffic_note": traffic_note,
            "fp64_dmma": dmma, "fp64_tensor_peak_tflops": fp64_peak,
        }
    else:
        traffic, traffic_note = measured_traffic("gemm_dmma_A")
        roof = dict(dmma)
        roof.update({
            "bound": "tensor",
            "peak_source": "FP64 tensor (DMMA) issue peak measured in THIS run by mogpe_probe_fp64_tensor_peak (independent "
                           f"mma.sync.m8n8k4.f64 chains on every SM; round-1 value {FP64_TENSOR_PEAK_FALLBACK}, cuBLAS DGEMM "
                           "35.4); MEASURED_PEAKS.json holds only HBM and bf16 figures",
            "gemm_ms_per_step": prof["gemm_ms"] / steps, "gemm_share_of_step": prof["gemm_ms"] / ms_prof,
            "profiled_pass_ms_per_step": ms_prof / steps,
            "step_model_tflops": step_flops(wl, B) * steps / (ms * 1e-3) * 1e-12,
            "traffic": traffic, "traffic_note": traffic_note})
    line = {
        "metric": "elbo_fwd_bwd_points_per_sec", "value": value, "unit": "points/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "steps_per_sec": steps / (ms * 1e-3),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg, "clocks": clocks,
        "e2e": {"value": pts / (ms_e2e * 1e-3), "unit": "points/s", "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": int(B * (wl["D"] + wl["F"]) * 8), "d2h_bytes_per_step": 32,
                "api": "mogpe_b200.keras.MixtureOfSVGPExperts.fit(X_host, Y_host, batch_size=B): per step H2D of the "
                       "minibatch, ELBO fwd+bwd, [NCCL all-reduce], device-side Adam, D2H of the result",
                "final_loss": float(hist["loss"][-1])},
        "gpu_launches": int(launches_timed),
        "roofline": roof,
        "elbo_last_step": float(elbo_val[0]),
    }
    if dp_parity is not None:
        line["dp_parity"] = dp_parity
    # HBM-bound streaming kernels, each timed alone with CUDA events in the profiled pass: algorithmic bytes
    # / duration against the measured copy bandwidth
    hbm = HBM_PEAK_GBS
    line["roofline"]["streaming"] = {
        k: {"us_per_step": 1e3 * ms_k / steps, "achieved": by / (ms_k * 1e-3) * 1e-9, "peak": hbm, "unit": "GB/s",
            "frac": by / (ms_k * 1e-3) * 1e-9 / hbm}
        for k, (ms_k, by) in stream_prof.items() if ms_k > 0}
    if world == 1 and not args.no_predict:
        line["predict"] = run_predict_modes(local_rank, fp64_peak=fp64_peak)
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = run_cpu_reference(wl, spec, 3, 1, min(B, 8192))
        line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
