"""Data-parallel path on real GPUs (skipped with fewer than 2 devices): two ranks, NCCL all-reduce inside the
library (overlapped with the Cholesky backward, and as a separate call), result equals the single-GPU full-batch
ELBO and gradients.  bench.py --gpus N runs the same check at the bench shape (`dp_parity` in its JSON line), which
is how the 8-GPU box exercises it."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = {
    # name: make_spec kwargs (N = global batch)
    "small": dict(seed=6, N=300, D=2, F=1, K=2, M=20, S=2, num_data=30000),
    # K = 3 Softmax gating, M = 512 and 2048 rows per rank: the INT8-sliced products of the backward are active
    "m512": dict(seed=8, N=4096, D=3, F=1, K=3, M=512, S=1, num_data=400000, separate_kernels=True),
}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, case, overlap):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mogpe_b200 import parallel
    from mogpe_b200.engine import Engine
    from tests._engine_helpers import block_from_gp
    from tests._specs import make_spec

    kw = CASES[case]
    spec, X, Y, eps = make_spec(**kw)
    N, S = kw["N"], kw["S"]
    experts = [block_from_gp(g, True) for g in spec["experts"]]
    eng = Engine(experts, block_from_gp(spec["gating"], False), flavour="keras", max_batch=N, device=rank)
    uid = parallel.DataParallelEngine.exchange_unique_id(Engine, rank)
    eng.init_data_parallel(rank, world, uid, overlap=overlap)
    dp = parallel.DataParallelEngine(eng, rank, 1)  # communicator already joined above (with the overlap mode under test)
    dp.world_size = world
    sl = parallel.shard_slice(N, rank, world)
    Xl, Yl = parallel.shard_batch(X, Y, rank, world)
    eps_u = eng.pack_eps_u(eps["u_experts"], None, S)
    eps_h = np.ascontiguousarray(eps["h"][:, sl])
    for rep in range(2):  # twice: the second call reuses scratch and the communication stream
        res = dp.step(Xl, Yl, N, spec["num_data"], "further_gating", S, readback=True, eps_u=eps_u, eps_h=eps_h)
    flat = eng.gbuf.numpy()
    assert res.elbo == flat[-4]
    np.save(os.path.join(out_dir, f"dp{rank}.npy"), flat)
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
@pytest.mark.parametrize("case", sorted(CASES))
def test_two_gpu_data_parallel_matches_single_gpu(tmp_path, case, overlap):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tests._engine_helpers import engine_from_spec, run_engine_elbo
    from tests._specs import make_spec

    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), case, overlap), nprocs=2, join=True)
    got0, got1 = np.load(tmp_path / "dp0.npy"), np.load(tmp_path / "dp1.npy")
    assert np.array_equal(got0, got1)  # every rank holds the same reduced buffer
    spec, X, Y, eps = make_spec(**CASES[case])
    eng = engine_from_spec(spec, max_batch=CASES[case]["N"])
    res = run_engine_elbo(eng, spec, X, Y, eps, device_inputs=True)  # device entry: results land behind the gradients
    want = eng.gbuf.numpy()
    assert want[-4] == res.elbo and res.elbo != 0.0
    assert abs(got0[-4] - want[-4]) <= 1e-10 * abs(want[-4]), (got0[-4:], want[-4:])
    np.testing.assert_allclose(got0[-4:], want[-4:], rtol=1e-10, atol=1e-9)
    # gradients: 1e-8 of the largest entry of the buffer section they belong to (FP64 atomics reorder sums)
    lo = eng.lo
    bounds = [lo.lengthscales, lo.variance, lo.Z, lo.q_mu, lo.q_sqrt, lo.mean_c, lo.noise, lo.total]
    for a, b in zip(bounds[:-1], bounds[1:]):
        scale = max(1e-300, np.max(np.abs(want[a:b])))
        assert np.max(np.abs(got0[a:b] - want[a:b])) <= 1e-8 * scale, (a, b)


def _rng_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mogpe_b200 import parallel
    from mogpe_b200.engine import Engine
    from tests._engine_helpers import block_from_gp
    from tests._specs import make_spec

    spec, X, Y, eps = make_spec(**CASES["small"])
    experts = [block_from_gp(g, True) for g in spec["experts"]]
    eng = Engine(experts, block_from_gp(spec["gating"], False), flavour="keras", max_batch=300, device=rank)
    eng.set_seed(99)  # the same seed on every rank
    eng.init_data_parallel(rank, world, parallel.DataParallelEngine.exchange_unique_id(Engine, rank))
    Xl, Yl = parallel.shard_batch(X[:299], Y[:299], rank, world)  # 150 + 149 rows: unequal shards
    out = {}
    for step in range(2):
        eng.elbo(Xl, Yl, "further_gating", 2, 100.0, want_grads=True, unpack=False, readback=False)
        out[f"u{step}"] = eng.debug("eps_u")
        out[f"h{step}"] = eng.debug("eps_h")
    np.savez(os.path.join(out_dir, f"rng{rank}.npz"), **out)
    dist.destroy_process_group()


def test_library_draws_in_data_parallel_mode(tmp_path):
    """Library-drawn noise with one seed for the job: the inducing-point draws are identical on every rank in every step
    (also with unequal shards), the per-point draws of the shards are different streams."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_rng_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rng0.npz"), np.load(tmp_path / "rng1.npz")
    for step in range(2):
        assert np.array_equal(a[f"u{step}"], b[f"u{step}"]) and np.any(a[f"u{step}"] != 0)
        n = min(a[f"h{step}"].size, b[f"h{step}"].size)
        assert not np.array_equal(a[f"h{step}"][:n], b[f"h{step}"][:n])
    assert not np.array_equal(a["u0"], a["u1"])  # and the shared stream advances between steps
