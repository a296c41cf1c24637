"""ELBO and gradient error of the INT8 step by the number of byte slices its forward / gradient-only products load, and of the
fast training mode (six everywhere), against the all-FP64 step:
C4 bench shape and the ill-conditioned stress model of tests/test_gpu_engine.py (512 inducing points in 2-D)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests._engine_helpers import engine_from_spec, engine_grads_named, run_engine_elbo
from tests._specs import make_spec
from tests.test_gpu_bench_shapes import _workload

def run(spec, X, Y, eps, nb, tag):
    eng = engine_from_spec(spec, max_batch=nb)
    eng.set_int8_backward(False)
    r0 = run_engine_elbo(eng, spec, X, Y, eps, device_inputs=True)
    ref, ref_elbo = engine_grads_named(r0, spec), r0.elbo
    eng.set_int8_backward(True)
    for fs, ns in ((8, 8), (8, 7), (7, 7), (8, 6)):
        eng.set_int8_forward_slices(fs)
        eng.set_int8_grad_slices(ns)
        r = run_engine_elbo(eng, spec, X, Y, eps, device_inputs=True)
        got = engine_grads_named(r, spec)
        worst_max, worst_l2 = 0.0, 0.0
        for k, g in ref.items():
            g = np.asarray(g, float); a = np.asarray(got[k], float)
            worst_max = max(worst_max, np.max(np.abs(a - g)) / max(np.max(np.abs(g)), 1e-300))
            worst_l2 = max(worst_l2, np.linalg.norm((a - g).ravel()) / max(np.linalg.norm(g.ravel()), 1e-300))
        print(f"{tag}: forward {fs} / gradient {ns} slices: ELBO rel {abs(r.elbo - ref_elbo) / abs(ref_elbo):.2e}, "
              f"worst gradient block max-norm rel {worst_max:.2e}, L2 rel {worst_l2:.2e}", flush=True)
    # fast training mode: six slices in the forward products too
    eng.set_int8_grad_slices(7)
    eng.set_int8_forward_slices(6)
    r6 = run_engine_elbo(eng, spec, X, Y, eps, device_inputs=True)
    got = engine_grads_named(r6, spec)
    worst_max = max(np.max(np.abs(np.asarray(got[k], float) - np.asarray(g, float))) / max(np.max(np.abs(np.asarray(g, float))), 1e-300)
                    for k, g in ref.items())
    print(f"{tag}: fast mode (six slices everywhere): ELBO rel {abs(r6.elbo - ref_elbo) / abs(ref_elbo):.2e}, "
          f"worst block max-norm rel {worst_max:.2e}", flush=True)
    eng.close()

wl, spec, X, Y, eps = _workload("C4")
run(spec, X, Y, eps, wl["batch"], "C4 bench shape")
for bound in ("further_gating", "further"):
    spec, X, Y, eps = make_spec(seed=17, N=2048, D=2, F=1, K=2, M=512, S=1, flavour="keras", bound=bound, num_data=20000)
    run(spec, X, Y, eps, 2048, f"ill-conditioned M=512 D=2 ({bound})")
