"""predict_y / predict_mixing_probs throughput (BASELINE config C5: K=8, G=8, D=4, M=1024, N up to 1e8).
Test points are generated on the device; outputs stay on the device; Softmax Monte-Carlo draws come from the library stream.
Usage: python tools/predict_bench.py [N] [M] [float64|tf32]
       python -m torch.distributed.run --nproc-per-node G ... tools/predict_bench.py N M precision   (N sharded over G GPUs,
       no collective on the data path: SURVEY 8e; the time is the max over ranks)"""
import os, sys
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
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lo, hi = (N * rank) // world, (N * (rank + 1)) // world   # contiguous shard of the test points
n = hi - lo
wl = dict(D=4, F=1, K=8, M=M, batch=65536, num_data=N, sep=False)
spec = bench.synth_model(wl, seed=4)
experts, gating = bench.blocks_from_spec(spec)
eng = Engine(experts, gating, flavour="keras", max_batch=65536, precision=precision, device=local)
eng.set_seed(100 + rank)
lib = _lib.load()
X = DeviceBuffer((n, 4), local, zero=False)
eng.fill_normal(X, seed=4 + rank)
R = 100  # gpflow's Softmax Monte-Carlo draws per point, generated on the device from the library stream
outs = {k: DeviceBuffer(s, local, zero=False) for k, s in (("y_mean", (n, 1)), ("y_var", (n, 1)), ("mix", (n, 8)))}


def run(count):
    _lib.check(lib.mogpe_predict(eng.h, X.ptr, count, eng.params.ptr, None, None, outs["mix"].ptr, outs["y_mean"].ptr,
                                 outs["y_var"].ptr, None, R, None), eng.h)


run(min(n, 65536 * 2))  # warm-up on two chunks
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(n); e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t[0])
if rank == 0:
    flops = 16 * 2.0 * M * M * N  # (P_e + P_g) * 2 M^2 per point (BASELINE.md work model; A and q_sqrt^T A, triangular)
    print(f"[{precision}, {world} GPU] predict_y + predict_mixing_probs: N={N} M={M} K=8: {ms:.1f} ms  {N / ms * 1e3:.3e} points/s  "
          f"{flops / ms * 1e-9:.1f} TFLOP/s algorithmic ({flops / ms * 1e-9 / (37.0 * world):.2f} of DMMA peak)", flush=True)
    print("y_mean[:3]", outs["y_mean"].numpy(count=3), "mix[0]", outs["mix"].numpy(count=8))
if world > 1:
    dist.destroy_process_group()
