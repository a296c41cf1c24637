"""GEMM tile-configuration sweep on the shapes of the C4 step (tuning tool, not part of the product)."""
import sys
sys.path.insert(0, ".")
import torch
from mogpe_b200 import _lib
lib = _lib.load()
M, B, Q = 512, 8192, 16
T = torch.tril(torch.randn(Q, M, M, dtype=torch.float64, device="cuda"))
R = torch.randn(Q, M, B, dtype=torch.float64, device="cuda")
R2 = torch.randn(Q, M, B, dtype=torch.float64, device="cuda")
O = torch.zeros(Q, M, B, dtype=torch.float64, device="cuda")
O2 = torch.zeros(Q, M, M, dtype=torch.float64, device="cuda")
def run(name, fn, flops):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"  {name:28s} {ms:7.3f} ms  {flops / ms * 1e-9:6.2f} TF/s (algorithmic)")
u = float(M) * M * B * Q
for cfg in (7, 11, 12, 13, 14, 15, 16, 17):
    print("cfg", cfg)
    g = lambda A, Bm, C, Mm, N, K, ta, tb, km, tril: _lib.check(lib.mogpe_debug_gemm(0, A.data_ptr(), Bm.data_ptr(), C.data_ptr(), Q, Mm, N, K, ta, tb, km, tril, 1.0, 0.0, cfg, None))
    run("NN left-lower  T @ R", lambda: g(T, R, O, M, B, M, 0, 0, 1, 0), u)
    run("TN left-upper  T^T @ R", lambda: g(T, R, O, M, B, M, 1, 0, 2, 0), u)
    run("NT tril(R @ R2^T)", lambda: g(R, R2, O2, M, M, B, 0, 1, 0, 1), u)
    run("NN full        T @ R", lambda: g(T, R, O, M, B, M, 0, 0, 0, 0), 2 * u)
    run("TN M^3 upper x lower", lambda: g(T, T, O2, M, M, M, 1, 0, 6, 0), float(Q) * M**3 * 2 / 3)
