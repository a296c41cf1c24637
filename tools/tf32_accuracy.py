"""Accuracy of the TF32 mode against the float64 engine on the test specs and on the bench workloads (GPU)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from tests._specs import make_spec
from tests._engine_helpers import engine_from_spec

def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))

for (K, F, sep, M, N, D) in [(2, 1, False, 25, 333, 2), (3, 1, False, 70, 300, 2), (2, 2, True, 150, 700, 2), (3, 2, False, 260, 520, 2),
                             (2, 1, False, 512, 2048, 4), (8, 1, False, 512, 2048, 4)]:
    spec, X, Y, eps = make_spec(seed=40 + K + F, N=N, D=D, F=F, K=K, M=M, flavour="legacy", separate_kernels=sep)
    e64 = engine_from_spec(spec, max_batch=1024)
    e32 = engine_from_spec(spec, max_batch=1024, precision="tf32")
    eps_mc = np.random.default_rng(3).standard_normal((17, N, K)) if K > 2 else None
    a, b = e64.predict(X, eps_mc=eps_mc), e32.predict(X, eps_mc=eps_mc)
    kuu = e64.debug_copy("Kuu") if hasattr(e64, "debug_copy") else None
    print(f"K={K} F={F} M={M} N={N} D={D}: " + "  ".join(f"{k} {rel(b[k], a[k]):.2e}" for k in ("f_mean", "f_var", "mix_probs", "y_mean", "y_var")),
          f" min f_var/max f_var {a['f_var'].min() / a['f_var'].max():.2e}")
    e64.close(); e32.close()
