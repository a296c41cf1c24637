"""End-to-end fit() time per step with and without CUDA-graph replay.  Usage: python tools/fit_bench.py [C1|C2|C3|C4] [steps]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
import bench
from mogpe_b200._model_core import Adam

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
wl = bench.WORKLOADS[name]
spec = bench.synth_model(wl)
B = wl["batch"]
X, Y = bench.synth_batch(wl, B * 4, seed=1)
X = np.concatenate([X] * ((steps + 3) // 4 + 1))[: steps * B]
Y = np.concatenate([Y] * ((steps + 3) // 4 + 1))[: steps * B]
for graphs in (0, 2, 0, 2):  # 0 = direct launches, 2 = always replay as a CUDA graph (1 = automatic choice)
    model = bench.build_public_model(spec, wl)
    model.set_seed(1)
    model.compile(optimizer=Adam(learning_rate=1e-3))
    model._core.use_graphs = graphs
    model.fit(X[: 3 * B], Y[: 3 * B], epochs=1, batch_size=B, shuffle=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.fit(X, Y, epochs=1, batch_size=B, shuffle=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name} graphs={graphs}: {1e3 * dt / steps:.3f} ms/step ({steps} steps, wall clock incl. fit() set-up and pull)", flush=True)
    del model
