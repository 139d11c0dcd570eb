"""CPU, world_size 2 over gloo: the sharding + single sum-reduce plumbing of the polychromatic driver.
The per-unit compute is replaced by a host function (the CUDA kernels need a GPU); what is under test
is the host logic: every unit is owned by exactly one rank, partial planes are reduced once, and the
total equals the serial weighted sum (prysm/polynomials/fitting.py:37)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_units, dst, out_q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from prysm_b200.polychromatic import sharded_incoherent_sum, shard_units
    rng = np.random.default_rng(0)
    planes = rng.random((n_units, 6, 5))
    weights = rng.random(n_units)
    mine = shard_units(n_units, rank, world)
    seen = []

    def unit(i, acc):
        seen.append(i)
        acc += torch.from_numpy(weights[i] * planes[i])

    total = sharded_incoherent_sum(n_units, unit, torch.zeros(6, 5, dtype=torch.float64), dst=dst)
    assert seen == mine
    out_q.put((rank, mine, total.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_units,dst', [(7, None), (7, 0), (2, None), (1, 0)])
def test_sharded_sum_two_ranks(n_units, dst):
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, dst, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    planes = rng.random((n_units, 6, 5))
    weights = rng.random(n_units)
    want = np.tensordot(planes, weights, axes=(0, 0))
    owned = sorted(i for _, mine, _ in res for i in mine)
    assert owned == list(range(n_units))            # every unit exactly once
    for rank, _, total in res:
        if dst is None or rank == dst:
            assert np.allclose(total, want, rtol=1e-13, atol=1e-15)


def test_shard_units_properties():
    from prysm_b200.polychromatic import shard_units
    for n in (0, 1, 7, 64):
        for w in (1, 2, 4, 8):
            parts = [shard_units(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_units(4, 2, 2)
