"""predict_y / predict_mixing_probs throughput (BASELINE config C5: K=8, G=8, D=4, M=1024, N up to 1e8).
Test points are generated on the device; outputs stay on the device.  Usage: python tools/predict_bench.py [N] [M] [float64|tf32]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import bench
from mogpe_b200 import _lib
from mogpe_b200.device import DeviceBuffer
from mogpe_b200.engine import Engine
import torch

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
precision = sys.argv[3] if len(sys.argv) > 3 else "float64"
wl = dict(D=4, F=1, K=8, M=M, batch=65536, num_data=N, sep=False)
spec = bench.synth_model(wl, seed=4)
experts, gating = bench.blocks_from_spec(spec)
eng = Engine(experts, gating, flavour="keras", max_batch=65536, precision=precision)
lib = _lib.load()
X = DeviceBuffer((N, 4), zero=False)
eng.fill_normal(X, seed=4)
R = 100  # gpflow's Softmax Monte-Carlo draws per point, generated on the device from the library stream
outs = {k: DeviceBuffer(s, zero=False) for k, s in (("y_mean", (N, 1)), ("y_var", (N, 1)), ("mix", (N, 8)))}
def run():
    _lib.check(lib.mogpe_predict(eng.h, X.ptr, N, eng.params.ptr, None, None, outs["mix"].ptr, outs["y_mean"].ptr, outs["y_var"].ptr,
                                 None, R, None), eng.h)
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
flops = 16 * 2.0 * M * M * N  # (P_e + P_g) * 2 M^2 per point (BASELINE.md work model; A and q_sqrt^T A, triangular)
print(f"[{precision}] predict_y + predict_mixing_probs: N={N} M={M} K=8: {ms:.1f} ms  {N / ms * 1e3:.3e} points/s  "
      f"{flops / ms * 1e-9:.1f} TFLOP/s algorithmic ({flops / ms * 1e-9 / 37.0:.2f} of DMMA peak)")
print("y_mean[:3]", outs["y_mean"].numpy(count=3), "mix[0]", outs["mix"].numpy(count=8))
