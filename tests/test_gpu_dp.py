"""Data-parallel path on real GPUs (skipped with fewer than 2 devices): two ranks, NCCL all-reduce inside the
library, result equals the single-GPU full-batch ELBO and gradients."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
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

    spec, X, Y, eps = make_spec(seed=6, N=300, D=2, F=1, K=2, M=20, S=2, num_data=30000)
    experts = [block_from_gp(g, True) for g in spec["experts"]]
    eng = Engine(experts, block_from_gp(spec["gating"], False), flavour="keras", max_batch=300, device=rank)
    dp = parallel.DataParallelEngine(eng, rank, world, parallel.DataParallelEngine.exchange_unique_id(Engine, rank))
    sl = parallel.shard_slice(300, rank, world)
    Xl, Yl = parallel.shard_batch(X, Y, rank, world)
    eps_u = eng.pack_eps_u(eps["u_experts"], None, 2)
    eps_h = np.ascontiguousarray(eps["h"][:, sl])
    dp.step(Xl, Yl, 300, spec["num_data"], "further_gating", 2, eps_u=eps_u, eps_h=eps_h)
    flat = eng.gbuf.numpy()
    if rank == 0:
        np.save(os.path.join(out_dir, "dp.npy"), flat)
    dist.destroy_process_group()


def test_two_gpu_data_parallel_matches_single_gpu(tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tests._engine_helpers import engine_from_spec, run_engine_elbo
    from tests._specs import make_spec

    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "dp.npy")
    spec, X, Y, eps = make_spec(seed=6, N=300, D=2, F=1, K=2, M=20, S=2, num_data=30000)
    eng = engine_from_spec(spec, max_batch=300)
    res = run_engine_elbo(eng, spec, X, Y, eps)
    want = eng.gbuf.numpy()
    assert abs(got[-4] - want[-4]) <= 1e-10 * abs(want[-4])
    np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-9)
