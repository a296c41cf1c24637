"""Gradient difference between the INT8 (integer-sliced) and the FP64 DMMA evaluation of Lbar / Sbar on the bench workloads."""
import sys
sys.path.insert(0, ".")
import numpy as np
import bench
from mogpe_b200.engine import Engine

for name in ("C3", "C4"):
    wl = bench.WORKLOADS[name]
    spec = bench.synth_model(wl)
    experts, gating = bench.blocks_from_spec(spec)
    B = wl["batch"]
    X, Y = bench.synth_batch(wl, B, seed=100)
    eng = Engine(experts, gating, flavour=bench.FLAVOUR, max_batch=B)
    out = {}
    for on in (True, False):
        eng.set_int8_backward(on)
        eng.set_seed(7)
        res = eng.elbo(X, Y, bench.BOUND, bench.S, wl["num_data"] / B, want_grads=True, unpack=False)
        out[on] = (res.elbo, eng.gbuf.numpy(count=eng.lo.total).copy())
    ga, gb = out[True][1], out[False][1]
    lo = eng.lo
    parts = {"lengthscales": (lo.lengthscales, lo.variance), "variance": (lo.variance, lo.Z), "Z": (lo.Z, lo.q_mu), "q_mu": (lo.q_mu, lo.q_sqrt),
             "q_sqrt": (lo.q_sqrt, lo.mean_c), "mean_c+noise": (lo.mean_c, lo.total)}
    print(name, "elbo", out[True][0], out[False][0])
    for k, (a, b) in parts.items():
        d = np.max(np.abs(ga[a:b] - gb[a:b])) / max(1e-300, np.max(np.abs(gb[a:b])))
        print(f"  {k:14s} max |int8 - fp64| / max |fp64| = {d:.2e}")
    eng.close()
