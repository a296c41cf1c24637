"""Run a few C4 (or other workload) ELBO fwd+bwd steps on one GPU: the target of ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from mogpe_b200.device import DeviceBuffer
from mogpe_b200.engine import Engine

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C4"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec = bench.synth_model(wl)
experts, gating = bench.blocks_from_spec(spec)
B = wl["batch"]
eng = Engine(experts, gating, flavour=bench.FLAVOUR, max_batch=B, device=0)
x, y = bench.synth_batch(wl, B, seed=100)
xd, yd = DeviceBuffer.from_numpy(x, 0), DeviceBuffer.from_numpy(y, 0)
for i in range(steps):
    eng.elbo(xd, yd, bench.BOUND, bench.S, wl["num_data"] / B, want_grads=True, unpack=False, readback=False)
print("elbo", eng.gbuf.numpy(offset=eng.lo.total, count=4))
