# cuBLAS DGEMM ceiling probe (anchors the FP64 roofline; MEASURED_PEAKS.json has no FP64 figure).
import torch, time, json
torch.backends.cuda.matmul.allow_tf32 = False
res = {}
for n in (2048, 4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(3): c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}"] = 2 * n**3 / best * 1e-9
    print(f"cuBLAS DGEMM n={n}: {res[f'dgemm_{n}']:.2f} TFLOP/s ({best:.3f} ms)")
# batched shapes like ours: [16, 512, 512] x [16, 512, 8192]
a = torch.randn(16, 512, 512, dtype=torch.float64, device="cuda"); b = torch.randn(16, 512, 8192, dtype=torch.float64, device="cuda")
for _ in range(3): c = torch.bmm(a, b)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): c = torch.bmm(a, b)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
res["dbmm_16x512x512x8192"] = 2 * 16 * 512 * 512 * 8192 / ms * 1e-9
print(f"cuBLAS batched DGEMM 16x[512x512]x[512x8192]: {res['dbmm_16x512x512x8192']:.2f} TFLOP/s ({ms:.3f} ms)")
# cholesky / triangular solve baselines
k = torch.randn(16, 512, 512, dtype=torch.float64, device="cuda"); k = k @ k.transpose(1, 2) + 512 * torch.eye(512, dtype=torch.float64, device="cuda")
for _ in range(2): L = torch.linalg.cholesky(k)
torch.cuda.synchronize(); e0.record(); L = torch.linalg.cholesky(k); e1.record(); torch.cuda.synchronize()
print(f"torch batched cholesky 16x512: {e0.elapsed_time(e1):.3f} ms")
for _ in range(2): x = torch.linalg.solve_triangular(L, b, upper=False)
torch.cuda.synchronize(); e0.record(); x = torch.linalg.solve_triangular(L, b, upper=False); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"torch batched trsm 16x512x8192: {ms:.3f} ms = {16*512*512*8192/ms*1e-9:.2f} TFLOP/s")
json.dump(res, open("gpurun_out/fp64_probe.json", "w"))
