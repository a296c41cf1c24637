"""world_size = 2 over gloo on the CPU: the host-side data-parallel logic (row sharding, global scale,
KL weight 1/world, one sum all-reduce of [grads | value]) reproduces the single-process ELBO and gradients.
The per-rank arithmetic is done by the oracle here (no GPU in this container); the NCCL path on real GPUs is
covered by bench.py --gpus N and tests/test_gpu_dp.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mogpe_b200 import parallel
from tests._specs import make_spec


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ref_torch as rt

    torch.set_num_threads(1)
    spec, X, Y, eps = make_spec(seed=5, N=41, D=2, F=1, K=2, M=6, S=2, num_data=4100)
    sl = parallel.shard_slice(X.shape[0], rank, world)
    Xl, Yl = parallel.shard_batch(X, Y, rank, world)
    local = dict(spec)
    local["num_data"] = None  # apply the GLOBAL scale by hand
    model = rt.MixtureOfSVGPExperts(local)
    named = model.named_variables()
    eps_l = {"u_experts": eps["u_experts"], "h": eps["h"][:, sl]}
    kl = model.gating.prior_kl() + sum(e.prior_kl() for e in model.experts)
    data_term = model.elbo(Xl, Yl, eps_l) + kl  # elbo with scale 1 = data - KL
    val = data_term * parallel.global_scale(spec["num_data"], X.shape[0]) - kl * parallel.kl_weight(world)
    grads = torch.autograd.grad(val, [v for _, v in named])
    flat = torch.cat([g.reshape(-1) for g in grads] + [val.detach().reshape(1)])
    dist.all_reduce(flat)  # the ONE collective of a step
    if rank == 0:
        np.save(os.path.join(out_dir, "dp.npy"), flat.numpy())
    dist.destroy_process_group()


def test_two_rank_data_parallel_matches_single_process(tmp_path):
    from oracle import ref_torch as rt

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "dp.npy")
    spec, X, Y, eps = make_spec(seed=5, N=41, D=2, F=1, K=2, M=6, S=2, num_data=4100)
    val, grads = rt.MixtureOfSVGPExperts(spec).elbo_and_grads(X, Y, eps)
    want = np.concatenate([np.asarray(g).reshape(-1) for g in grads.values()] + [[val]])
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10)


def test_shard_slices_partition_the_batch():
    for n in (1, 7, 64, 65, 1000):
        for world in (1, 2, 3, 8):
            rows = np.concatenate([np.arange(n)[parallel.shard_slice(n, r, world)] for r in range(world)])
            assert np.array_equal(rows, np.arange(n))
            sizes = [len(range(*parallel.shard_slice(n, r, world).indices(n))) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_slice(10, 2, 2)
    assert parallel.global_scale(None, 64) == 1.0 and parallel.global_scale(1000, 50) == 20.0
