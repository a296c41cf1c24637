"""Time the six INT8 products of one C4 step under an epilogue tuning switch (MOGPE_I8_DBG, see gemm_i8.cuh Params::dbg)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mogpe_b200.device import DeviceBuffer
from mogpe_b200.engine import Engine

wl = bench.WORKLOADS["C4"]
spec = bench.synth_model(wl)
experts, gating = bench.blocks_from_spec(spec)
B = wl["batch"]
eng = Engine(experts, gating, flavour=bench.FLAVOUR, max_batch=B, device=0)
x, y = bench.synth_batch(wl, B, seed=100)
xd, yd = DeviceBuffer.from_numpy(x, 0), DeviceBuffer.from_numpy(y, 0)
for i in range(2):
    eng.elbo(xd, yd, bench.BOUND, bench.S, wl["num_data"] / B, want_grads=True, unpack=False, readback=False)
eng.profile(True)
print("---- timed step, MOGPE_I8_DBG =", os.environ.get("MOGPE_I8_DBG"), file=sys.stderr, flush=True)
eng.elbo(xd, yd, bench.BOUND, bench.S, wl["num_data"] / B, want_grads=True, unpack=False, readback=False)
eng.profile(False)
